/* TEST INFRASTRUCTURE — not part of the product.
 *
 * Headless harness around THIS REPO's drop-in (blinky_b200/host/fisheye_b200.c),
 * compiled against the reference engine's own headers and the same engine stubs
 * as ref_harness.c.  It proves the drop-in claim: the C file builds where
 * engine/NQ/fisheye.c builds and F_Init / F_RenderView / F_WriteConfig produce the
 * same console text, config text and frame bytes as the reference's.
 * Exports `dropin_*` (ctypes, tests only).
 */
#define _GNU_SOURCE
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "bspfile.h"
#include "client.h"
#include "cmd.h"
#include "console.h"
#include "cvar.h"
#include "draw.h"
#include "fisheye.h"
#include "host.h"
#include "mathlib.h"
#include "quakedef.h"
#include "r_local.h"
#include "screen.h"
#include "sys.h"
#include "view.h"

#include "blinky_b200.h"

#define MAX_PLATES BLINKY_MAX_PLATES
#include "engine_stubs.inc"

blinky_ctx *F_B200_Context(void);
int F_B200_Building(void);
void F_B200_WaitBuild(void);
int F_B200_ShownDisplay(int *display, int *numplates);

static int harness_displayed_plate(int k, int *platesize)
{
    int display[BLINKY_MAX_PLATES], seen = 0, numplates = 0;
    blinky_ctx *c = F_B200_Context();
    if (!c) return -1;
    *platesize = scr_vrect.width < scr_vrect.height ? scr_vrect.width : scr_vrect.height;
    F_B200_ShownDisplay(display, &numplates); /* the plates of the lensmap that is on screen */
    for (int i = 0; i < numplates; i++) {
        if (display[i]) {
            if (seen == k) return i;
            seen++;
        }
    }
    return -1;
}

static byte *g_vidbuf;
static int g_inited;

int dropin_init(const char *basedir, const unsigned char *palette768)
{
    if (g_inited) return -1;
    snprintf(com_basedir, sizeof com_basedir, "%s", basedir);
    memcpy(g_palette, palette768, 768);
    host_basepal = g_palette;
    g_ncmds = 0;
    F_Init();
    g_inited = 1;
    return F_B200_Context() ? 0 : -2;
}

void dropin_shutdown(void)
{
    if (g_inited) F_Shutdown();
    g_inited = 0;
}

void dropin_set_gamedir(const char *dir) { snprintf(com_gamedir, sizeof com_gamedir, "%s", dir); }
void dropin_command(const char *text) { Cmd_ExecuteString(text, src_command); }
const char *dropin_log(void) { return g_log; }
void dropin_log_clear(void) { g_log_len = 0; g_log[0] = 0; }
const char *dropin_unhandled_commands(void) { return g_unhandled; }
int dropin_fisheye_enabled(void) { return fisheye_enabled; }
extern double fisheye_plate_fov;
double dropin_plate_fov(void) { return fisheye_plate_fov; }
int dropin_recalc_refdef(void) { int r = vid.recalc_refdef; vid.recalc_refdef = 0; return r; }

int dropin_set_screen(int w, int h, int rowbytes, int vx, int vy, int vw, int vh)
{
    free(g_vidbuf);
    g_vidbuf = (byte *)calloc((size_t)rowbytes * h, 1);
    if (!g_vidbuf) return -1;
    vid.buffer = g_vidbuf;
    vid.width = w;
    vid.height = h;
    vid.rowbytes = rowbytes;
    vid.aspect = (float)w / (float)h;
    scr_vrect.x = vx;
    scr_vrect.y = vy;
    scr_vrect.width = vw;
    scr_vrect.height = vh;
    return 0;
}

/* one F_RenderView call; the lensmap rebuild it may start runs on the drop-in's worker thread */
int dropin_frame_nowait(const unsigned char *faces, const unsigned char *background, unsigned char *out)
{
    g_faces = faces;
    g_background = background;
    g_render_calls = 0;
    F_RenderView();
    if (out) memcpy(out, vid.buffer, (size_t)vid.rowbytes * vid.height);
    g_faces = NULL;
    g_background = NULL;
    return g_render_calls;
}
int dropin_building(void) { return F_B200_Building(); }

/* the frame a viewer sees once a pending rebuild has finished: F_RenderView (which starts the rebuild
 * if the console changed anything), wait for the worker, F_RenderView again */
int dropin_frame(const unsigned char *faces, const unsigned char *background, unsigned char *out)
{
    g_faces = faces;
    g_background = background;
    g_render_calls = 0;
    F_RenderView();
    if (F_B200_Building()) {
        F_B200_WaitBuild();
        g_render_calls = 0;
        F_RenderView();
    }
    if (out) memcpy(out, vid.buffer, (size_t)vid.rowbytes * vid.height);
    g_faces = NULL;
    g_background = NULL;
    return g_render_calls;
}

int dropin_write_config(const char *path)
{
    FILE *f = fopen(path, "w");
    if (!f) return -1;
    F_WriteConfig(f);
    fclose(f);
    return 0;
}
