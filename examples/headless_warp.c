/* headless_warp.c — the C ABI without an engine: build a lensmap, warp a batch of frames.
 *
 *   gcc -O2 -Iinclude examples/headless_warp.c -Lblinky_b200 -lblinky_b200 -Wl,-rpath,$PWD/blinky_b200 -o headless_warp
 *   ./headless_warp [device [lens [width height platesize [frames]]]]
 *
 * Scripts are read from $BLINKY_BASEDIR/lua-scripts/{globes,lenses} (default: ./blinky_b200, i.e. run
 * it from the repository root; point it at the reference's game/ directory to use its scripts).
 *
 * device < 0 gives a host-only context: scripts, console and the lensmap build work (interpreter),
 * the warp refuses with BLINKY_E_NODEVICE — there is no CPU fallback for the hot path.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "blinky_b200.h"

static void to_stdout(const char *text, void *user)
{
    (void)user;
    fputs(text, stdout); /* what the engine would route to Con_Printf */
}

#define CHECK(call)                                                                      \
    do {                                                                                 \
        int rc_ = (call);                                                                \
        if (rc_ != BLINKY_OK) {                                                          \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc_, blinky_last_error(ctx));       \
            blinky_destroy(ctx);                                                         \
            return rc_ == BLINKY_E_NODEVICE ? 3 : 1;                                     \
        }                                                                                \
    } while (0)

int main(int argc, char **argv)
{
    int device = argc > 1 ? atoi(argv[1]) : 0;
    const char *lens = argc > 2 ? argv[2] : "panini";
    int w = argc > 5 ? atoi(argv[3]) : 1920, h = argc > 5 ? atoi(argv[4]) : 1080, ps = argc > 5 ? atoi(argv[5]) : 1024;
    int frames = argc > 6 ? atoi(argv[6]) : 8;
    blinky_ctx *ctx = NULL;
    unsigned char palette[768];
    char cmd[128];
    int i, rc;

    rc = blinky_create(device, &ctx);
    if (rc != BLINKY_OK) {
        fprintf(stderr, "blinky_create(%d): %s\n", device, ctx ? blinky_last_error(ctx) : "out of memory");
        if (ctx) blinky_destroy(ctx);
        return 1;
    }
    blinky_set_print_callback(ctx, to_stdout, NULL);
    CHECK(blinky_set_basedir(ctx, getenv("BLINKY_BASEDIR") ? getenv("BLINKY_BASEDIR") : "blinky_b200"));
    for (i = 0; i < 768; i++) palette[i] = (unsigned char)(i * 37 + 11); /* any 256-colour palette */
    CHECK(blinky_set_palette(ctx, palette));
    blinky_command(ctx, "f_globe cube");
    snprintf(cmd, sizeof cmd, "f_lens %s", lens);
    blinky_command(ctx, cmd); /* runs the script and its onload zoom command */

    CHECK(blinky_build_lensmap(ctx, w, h, ps, 0)); /* 0: GPU build when there is one, else the interpreter */
    printf("lensmap %dx%d over %d plates of %d^2: %lld mapped pixels\n  built by: %s\n  layout: %s\n", blinky_width(ctx), blinky_height(ctx),
           blinky_numplates(ctx), blinky_platesize(ctx), (long long)blinky_mapped_pixels(ctx), blinky_build_info(ctx), blinky_plan_summary(ctx));

    {
        size_t face_stride = (size_t)BLINKY_MAX_PLATES * ps * ps, frame_bytes = (size_t)w * h;
        unsigned char *faces = NULL, *screen = NULL;
        size_t k;
        rc = blinky_alloc_pinned(ctx, face_stride * frames, (void **)&faces);
        if (rc == BLINKY_OK) rc = blinky_alloc_pinned(ctx, frame_bytes * frames, (void **)&screen);
        if (rc != BLINKY_OK) {
            fprintf(stderr, "no pinned memory (%d): %s\n", rc, blinky_last_error(ctx));
            blinky_destroy(ctx);
            return rc == BLINKY_E_NODEVICE ? 3 : 1;
        }
        for (k = 0; k < face_stride * frames; k++) faces[k] = (unsigned char)(k * 2654435761u >> 13); /* stand-in for rendered plates */
        CHECK(blinky_warp_host(ctx, faces, face_stride, screen, frame_bytes, w, 0, 0, frames, /*keep_unmapped*/ 0));
        printf("warped %d frames; first bytes of frame 0: %u %u %u %u; kernel: %s\n", frames, screen[0], screen[1], screen[2], screen[3],
               blinky_last_kernel(ctx));
        blinky_free_pinned(ctx, faces);
        blinky_free_pinned(ctx, screen);
    }
    blinky_destroy(ctx);
    return 0;
}
