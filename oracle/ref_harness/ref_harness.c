/* TEST INFRASTRUCTURE — not part of the product.
 *
 * Headless harness around the UNMODIFIED reference fisheye layer.  The
 * reference source is pulled in where it lies (`#include "fisheye.c"` resolves
 * through -I/root/reference/engine/NQ, see ../Makefile); nothing is copied into
 * this repository.  Including it in this translation unit gives the harness
 * access to the file-static `lens`, `globe`, `zoom`, `rubix` structures
 * (engine/NQ/fisheye.c:334-528) so the lensmap can be exported as indices.
 *
 * What is stubbed: the TyrQuake engine symbols fisheye.c imports
 * (SURVEY.md section 8b "Imports"): console/cmd, zone, shell completion, the
 * software renderer entry points and the video globals.  `R_RenderView` paints
 * the caller-supplied synthetic globe face, `Draw_TileClear` paints the
 * caller-supplied background.
 *
 * Exports (all `ref_*`, plain C ABI, loaded with ctypes from tests/ and
 * bench.py's CPU-baseline leg only).
 */
#define _GNU_SOURCE
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "fisheye.c" /* the reference, in place (engine/NQ/fisheye.c) */

/* ------------------------------------------------------------------------ */
/* engine data the fisheye layer reads                                       */
/* ------------------------------------------------------------------------ */
viddef_t vid;
vrect_t scr_vrect;
refdef_t r_refdef;
int sb_lines = 0;
byte *host_basepal;
char com_basedir[MAX_OSPATH];
cmd_source_t cmd_source;
static short LittleShort_impl(short l) { return l; }
short (*LittleShort)(short l) = LittleShort_impl;

static byte g_palette[768];

/* ------------------------------------------------------------------------ */
/* console                                                                   */
/* ------------------------------------------------------------------------ */
static char g_log[1 << 16];
static size_t g_log_len;

void Con_Printf(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    if (g_log_len < sizeof(g_log) - 1) {
        int n = vsnprintf(g_log + g_log_len, sizeof(g_log) - g_log_len, fmt, ap);
        if (n > 0) {
            g_log_len += (size_t)n;
            if (g_log_len > sizeof(g_log) - 1) g_log_len = sizeof(g_log) - 1;
        }
    }
    va_end(ap);
}

void Sys_Error(const char *error, ...)
{
    va_list ap;
    va_start(ap, error);
    vfprintf(stderr, error, ap);
    va_end(ap);
    fputc('\n', stderr);
    abort();
}

/* ------------------------------------------------------------------------ */
/* cmd                                                                       */
/* ------------------------------------------------------------------------ */
#define MAX_CMDS 64
static struct { const char *name; xcommand_t fn; } g_cmds[MAX_CMDS];
static int g_ncmds;
static int g_argc;
static char g_argv_store[16][256];
static char g_unhandled[4096]; /* commands executed but not registered (e.g. "bind ...") */

void Cmd_AddCommand(const char *cmd_name, xcommand_t function)
{
    if (g_ncmds < MAX_CMDS) {
        g_cmds[g_ncmds].name = cmd_name;
        g_cmds[g_ncmds].fn = function;
        g_ncmds++;
    }
}
void Cmd_SetCompletion(const char *cmd_name, cmd_arg_f completion) { (void)cmd_name; (void)completion; }
int Cmd_Argc(void) { return g_argc; }
const char *Cmd_Argv(int arg) { return (arg >= 0 && arg < g_argc) ? g_argv_store[arg] : ""; }

/* Tokenise like Quake's COM_Parse: whitespace separated, "quoted strings" kept whole. */
static void tokenize(const char *text)
{
    g_argc = 0;
    const char *p = text;
    while (*p && g_argc < 16) {
        while (*p == ' ' || *p == '\t') p++;
        if (!*p || *p == '\n' || *p == ';') break;
        char *out = g_argv_store[g_argc];
        size_t n = 0;
        if (*p == '"') {
            p++;
            while (*p && *p != '"' && n < 255) out[n++] = *p++;
            if (*p == '"') p++;
        } else {
            while (*p && *p != ' ' && *p != '\t' && *p != '\n' && *p != ';' && n < 255) out[n++] = *p++;
        }
        out[n] = 0;
        g_argc++;
    }
}

void Cmd_ExecuteString(const char *text, cmd_source_t src)
{
    (void)src;
    tokenize(text);
    if (!g_argc) return;
    for (int i = 0; i < g_ncmds; i++) {
        if (!strcasecmp(g_cmds[i].name, g_argv_store[0])) {
            g_cmds[i].fn();
            return;
        }
    }
    size_t l = strlen(g_unhandled);
    snprintf(g_unhandled + l, sizeof(g_unhandled) - l, "%s\n", text);
}

int Q_atoi(const char *str) { return atoi(str); }
float Q_atof(const char *str) { return (float)atof(str); }

/* ------------------------------------------------------------------------ */
/* zone / shell / fs                                                         */
/* ------------------------------------------------------------------------ */
void *Z_Malloc(int size) { return calloc(1, (size_t)size); }
static void *g_temp;
void *Hunk_TempAlloc(int size)
{
    free(g_temp);
    g_temp = calloc(1, (size_t)size);
    return g_temp;
}
void STree_AllocInit(void) {}
void COM_ScanDir(struct stree_root *root, const char *path, const char *pfx, const char *ext, qboolean stripext)
{
    (void)root; (void)path; (void)pfx; (void)ext; (void)stripext;
}
static char g_write_dir[MAX_OSPATH] = ".";
void COM_WriteFile(const char *filename, const void *data, int len)
{
    char path[MAX_OSPATH * 2];
    snprintf(path, sizeof path, "%s/%s", g_write_dir, filename);
    FILE *f = fopen(path, "wb");
    if (!f) return;
    fwrite(data, 1, (size_t)len, f);
    fclose(f);
}

/* ------------------------------------------------------------------------ */
/* renderer                                                                  */
/* ------------------------------------------------------------------------ */
static const byte *g_faces;      /* [numplates][ps][ps] supplied by the caller */
static const byte *g_background; /* [vid.height][vid.width] supplied by the caller */
static int g_render_calls;
static int g_rendered_plate[MAX_PLATES];

void D_EnableBackBufferAccess(void) {}
void D_DisableBackBufferAccess(void) {}
void R_PushDlights(void) {}
void R_SetVrect(const vrect_t *pvrectin, vrect_t *pvrect, int lineadj)
{
    /* the harness keeps scr_vrect == full screen; see ref_set_screen */
    (void)pvrectin; (void)pvrect; (void)lineadj;
}
void R_ViewChanged(vrect_t *pvrect, int lineadj, float aspect) { (void)pvrect; (void)lineadj; (void)aspect; }

/* F_RenderView calls render_plate() only for plates with display != 0, in
 * index order (fisheye.c:764-794), so the k-th call is the k-th displayed plate. */
void R_RenderView(void)
{
    int k = g_render_calls++;
    int plate = -1, seen = 0;
    for (int i = 0; i < globe.numplates; i++) {
        if (globe.plates[i].display) {
            if (seen == k) { plate = i; break; }
            seen++;
        }
    }
    if (plate < 0) return;
    if (k < MAX_PLATES) g_rendered_plate[k] = plate;
    if (!g_faces) return;
    int ps = globe.platesize;
    const byte *src = g_faces + (size_t)plate * ps * ps;
    for (int y = 0; y < ps; y++)
        memcpy(vid.buffer + scr_vrect.x + (size_t)(y + scr_vrect.y) * vid.rowbytes, src + (size_t)y * ps, (size_t)ps);
}

void Draw_TileClear(int x, int y, int w, int h)
{
    for (int row = y; row < y + h; row++) {
        if (g_background)
            memcpy(vid.buffer + x + (size_t)row * vid.rowbytes, g_background + x + (size_t)row * vid.width, (size_t)w);
        else
            memset(vid.buffer + x + (size_t)row * vid.rowbytes, 0, (size_t)w);
    }
}

/* ------------------------------------------------------------------------ */
/* exported harness API                                                      */
/* ------------------------------------------------------------------------ */
static byte *g_vidbuf;
static int g_inited;
static int g_alloc_w = -1, g_alloc_h = -1, g_alloc_ps = -1; /* sizes of lens/globe buffers */

/* basedir: directory containing lua-scripts/{globes,lenses}.  palette: 768 B RGB. */
int ref_init(const char *basedir, const unsigned char *palette768)
{
    if (g_inited) return -1;
    snprintf(com_basedir, sizeof com_basedir, "%s", basedir);
    memcpy(g_palette, palette768, 768);
    host_basepal = g_palette;
    g_ncmds = 0;
    F_Init();
    /* one-shot builds: never yield to the per-frame time slice (fisheye.c:819-826) */
    lens_builder.seconds_per_frame = 1.0e30f;
    g_inited = 1;
    return 0;
}

void ref_shutdown(void)
{
    if (!g_inited) return;
    F_Shutdown();
    g_inited = 0;
}

void ref_set_basedir(const char *basedir) { snprintf(com_basedir, sizeof com_basedir, "%s", basedir); }
void ref_set_write_dir(const char *dir) { snprintf(g_write_dir, sizeof g_write_dir, "%s", dir); }

void ref_command(const char *text) { Cmd_ExecuteString(text, src_command); }

const char *ref_log(void) { return g_log; }
void ref_log_clear(void) { g_log_len = 0; g_log[0] = 0; }
const char *ref_unhandled_commands(void) { return g_unhandled; }

/* Screen of w x h with the view rectangle (vx,vy,vw,vh) inside it and the given
 * row pitch, like vid/scr_vrect in the engine (include/vid.h:39-57). */
int ref_set_screen(int w, int h, int rowbytes, int vx, int vy, int vw, int vh)
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

/* One full engine frame through the real F_RenderView (platesize = min(w,h),
 * fisheye.c:707).  faces: [numplates][ps][ps]; background: [h][w]; out: [h][rowbytes]. */
int ref_frame(const unsigned char *faces, const unsigned char *background, unsigned char *out)
{
    g_faces = faces;
    g_background = background;
    g_render_calls = 0;
    /* F_RenderView only reallocates when the size differs from ITS previous call
     * (static pwidth/pheight, fisheye.c:700-727); if ref_build() resized the
     * buffers in between, put back buffers of the size F_RenderView assumes. */
    {
        int w = scr_vrect.width, h = scr_vrect.height, ps = w < h ? w : h;
        if (g_alloc_w != w || g_alloc_h != h || g_alloc_ps != ps) {
            free(globe.pixels);
            free(lens.pixels);
            free(lens.pixel_tints);
            globe.pixels = (byte *)malloc((size_t)ps * ps * MAX_PLATES);
            lens.pixels = (byte **)malloc((size_t)w * h * sizeof(byte *));
            lens.pixel_tints = (byte *)malloc((size_t)w * h);
            g_alloc_w = w; g_alloc_h = h; g_alloc_ps = ps;
            lens.changed = true;
        }
    }
    F_RenderView();
    while (lens_builder.working) resume_lensmap();
    if (out) memcpy(out, vid.buffer, (size_t)vid.rowbytes * vid.height);
    g_faces = NULL;
    g_background = NULL;
    return g_render_calls;
}

/* Lensmap (re)build with the plate size DECOUPLED from the screen size (the
 * BASELINE configs use 256/1024/2048-pixel faces).  This replays exactly the
 * rebuild branch of F_RenderView (fisheye.c:704-743) with `platesize` chosen by
 * the caller instead of min(w,h); every function called is the reference's. */
int ref_build(int width_px, int height_px, int platesize)
{
    lens.width_px = width_px;
    lens.height_px = height_px;
    globe.platesize = platesize;
    int area = width_px * height_px;

    if (globe.pixels) free(globe.pixels);
    if (lens.pixels) free(lens.pixels);
    if (lens.pixel_tints) free(lens.pixel_tints);
    globe.pixels = (byte *)malloc((size_t)platesize * platesize * MAX_PLATES);
    lens.pixels = (byte **)malloc((size_t)area * sizeof(byte *));
    lens.pixel_tints = (byte *)malloc((size_t)area);
    if (!globe.pixels || !lens.pixels || !lens.pixel_tints) return -1;
    g_alloc_w = width_px; g_alloc_h = height_px; g_alloc_ps = platesize;

    memset(lens.pixels, 0, (size_t)area * sizeof(byte *));
    memset(lens.pixel_tints, 255, (size_t)area);
    lens.valid = LUA_load_lens();
    if (!lens.valid) {
        strcpy(lens.name, "");
        Con_Printf("not a valid lens\n");
    }
    create_lensmap();
    while (lens_builder.working) resume_lensmap();
    lens.changed = globe.changed = zoom.changed = false;
    return (lens.valid && globe.valid) ? 0 : 1;
}

/* idx[i] = offset of the source texel inside globe.pixels, or -1 when unmapped */
void ref_get_lensmap(int *idx, unsigned char *tint)
{
    int area = lens.width_px * lens.height_px;
    for (int i = 0; i < area; i++) {
        idx[i] = lens.pixels[i] ? (int)(lens.pixels[i] - globe.pixels) : -1;
        if (tint) tint[i] = lens.pixel_tints[i];
    }
}

void ref_get_palmaps(unsigned char *out /* [6][256] */)
{
    for (int j = 0; j < MAX_PLATES; j++) memcpy(out + j * 256, globe.plates[j].palette, 256);
}

int ref_numplates(void) { return globe.numplates; }
int ref_platesize(void) { return globe.platesize; }
double ref_scale(void) { return lens.scale; }
int ref_map_type(void) { return (int)lens.map_type; }
int ref_lens_valid(void) { return lens.valid; }
int ref_globe_valid(void) { return globe.valid; }
int ref_rubix_enabled(void) { return rubix.enabled; }
double ref_lens_width(void) { return lens.width; }
double ref_lens_height(void) { return lens.height; }
int ref_max_fov(void) { return zoom.max_fov; }
int ref_max_vfov(void) { return zoom.max_vfov; }
int ref_fisheye_enabled(void) { return fisheye_enabled; }
double ref_plate_fov(void) { return fisheye_plate_fov; }

void ref_get_display(int *out /* [6] */)
{
    for (int i = 0; i < MAX_PLATES; i++) out[i] = i < globe.numplates ? globe.plates[i].display : 0;
}

/* out: per plate 11 floats: forward[3] right[3] up[3] fov dist */
void ref_get_plates(float *out)
{
    for (int i = 0; i < globe.numplates; i++) {
        float *o = out + i * 11;
        memcpy(o, globe.plates[i].forward, 12);
        memcpy(o + 3, globe.plates[i].right, 12);
        memcpy(o + 6, globe.plates[i].up, 12);
        o[9] = globe.plates[i].fov;
        o[10] = globe.plates[i].dist;
    }
}

/* The hot loop alone: render_lensmap() (fisheye.c:2406-2424) over caller faces.
 * faces: [6][ps][ps] (copied into globe.pixels), background/out: [vid.height][vid.rowbytes]. */
void ref_render(const unsigned char *faces, int numfaces, const unsigned char *background, unsigned char *out)
{
    size_t ps2 = (size_t)globe.platesize * globe.platesize;
    memcpy(globe.pixels, faces, ps2 * (size_t)numfaces);
    if (background) memcpy(vid.buffer, background, (size_t)vid.rowbytes * vid.height);
    render_lensmap();
    if (out) memcpy(out, vid.buffer, (size_t)vid.rowbytes * vid.height);
}

/* Times `reps` calls of render_lensmap() on the current lensmap/faces and
 * returns the best single-call wall time in seconds (CLOCK_MONOTONIC). */
double ref_time_render(int reps, double *total_seconds)
{
    double best = 1e30, total = 0;
    for (int r = 0; r < reps; r++) {
        struct timespec a, b;
        clock_gettime(CLOCK_MONOTONIC, &a);
        render_lensmap();
        clock_gettime(CLOCK_MONOTONIC, &b);
        double dt = (double)(b.tv_sec - a.tv_sec) + 1e-9 * (double)(b.tv_nsec - a.tv_nsec);
        if (dt < best) best = dt;
        total += dt;
    }
    if (total_seconds) *total_seconds = total;
    return best;
}

int ref_write_config(const char *path)
{
    FILE *f = fopen(path, "w");
    if (!f) return -1;
    F_WriteConfig(f);
    fclose(f);
    return 0;
}

/* direct probes of the reference's Lua<->C bridge, for interpreter parity tests */
int ref_lens_inverse(double x, double y, float *ray)
{
    vec3_t r = {0, 0, 0};
    if (lua_refs.lens_inverse == -1) return -2;
    int st = LUAtoC_lens_inverse(x, y, r);
    ray[0] = r[0]; ray[1] = r[1]; ray[2] = r[2];
    return st;
}

int ref_lens_forward(const float *ray, double *x, double *y)
{
    vec3_t r = {ray[0], ray[1], ray[2]};
    if (lua_refs.lens_forward == -1) return -2;
    return LUAtoC_lens_forward(r, x, y);
}

int ref_ray_to_plate(const float *ray, double *u, double *v)
{
    vec3_t r = {ray[0], ray[1], ray[2]};
    int p = ray_to_plate_index(r);
    if (p < 0) return -1;
    if (!ray_to_plate_uv(p, r, u, v)) return -2 - p;
    return p;
}
