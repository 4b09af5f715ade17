/* probe_loop.c — the reference's interactive pipeline (src/main.rs + the probe closure of src/pipeline.rs:67-184) against the C ABI:
 * keyboard bytes -> UserCommand (src/raw_mode_guard.rs:65-101) -> TrackerContext::handle_command, and one vt_probe_frame per buffer
 * (convert + state machine + tracker + overlays + HUD, in place on the frame).
 *
 *   gcc -std=c99 -Iinclude examples/probe_loop.c -Lgstreamer_vit_tracker_b200 -lvittrack_b200 -o probe_loop
 *   ./probe_loop model.vtw frames.nv12 1920 1080 "dd  " > overlaid.nv12
 *
 * The key string plays the role of the TTY: one byte is consumed per frame ("d" moves the selection cursor right, the first space
 * confirms the first corner, the second one the opposite corner and starts tracking; 'q' quits).  Needs a B200. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "vt_tracker.h"

int main(int argc, char** argv) {
    if (argc != 6) {
        fprintf(stderr, "usage: %s model.vtw frames.nv12 width height keys > overlaid.nv12\n", argv[0]);
        return 2;
    }
    vt_config cfg;
    vt_config_default(&cfg);
    cfg.weights_path = argv[1];
    cfg.width = atoi(argv[3]), cfg.height = atoi(argv[4]);
    const size_t frame_bytes = (size_t)cfg.width * (size_t)cfg.height * 3 / 2;
    vt_context* ctx = NULL;
    if (vt_context_create(&cfg, &ctx) != VT_OK) {
        fprintf(stderr, "vt_context_create: %s\n", vt_last_error());
        return 1;
    }
    void* pinned = NULL;
    FILE* f = fopen(argv[2], "rb");
    if (!f || vt_alloc_pinned(frame_bytes, &pinned) != VT_OK) {
        fprintf(stderr, "cannot open %s / allocate the frame\n", argv[2]);
        return 1;
    }
    uint8_t* frame = (uint8_t*)pinned;
    const char* keys = argv[5];
    size_t nkeys = strlen(keys), k = 0;
    long n = 0;
    while (fread(frame, 1, frame_bytes, f) == frame_bytes) {
        if (k < nkeys) { /* ≙ the keyboard thread of src/main.rs: byte -> command -> channel -> handle_command */
            int32_t cmd = 0, fast = 0;
            if (vt_command_from_key((uint8_t)keys[k++], &cmd, &fast)) {
                if (cmd == VT_CMD_QUIT) break;
                vt_context_handle_command(ctx, cmd, fast);
            }
        }
        if (vt_probe_frame(ctx, frame, frame_bytes, NULL) != VT_OK) { /* the probe itself always returns Ok (src/pipeline.rs:183) */
            fprintf(stderr, "vt_probe_frame: %s\n", vt_last_error());
            return 1;
        }
        fprintf(stderr, "frame %ld: %s score %.2f\n", n++, vt_context_state_name(ctx), vt_context_current_score(ctx));
        fwrite(frame, 1, frame_bytes, stdout);
    }
    fclose(f);
    vt_free_pinned(pinned);
    vt_context_destroy(ctx);
    return 0;
}
