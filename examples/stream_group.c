/* stream_group.c — n independent video streams on one GPU, stepped together (≙ n pipelines of src/pipeline.rs:55, one TrackerContext
 * each, whose probes fire in the same frame period): every stream keeps its own pinned frame, search window, device-side rect_last and
 * box overlay; all of them go through ONE batched ViT forward per step (vt_tracker_update_streams).
 *
 *   gcc -std=c99 -Iinclude examples/stream_group.c -Lgstreamer_vit_tracker_b200 -lvittrack_b200 -o stream_group
 *   ./stream_group model.vtw 1920 1080 x y w h  a.nv12 b.nv12 [c.nv12 ...]      (2..16 files of tightly packed NV12 frames)
 *
 * The box is the target in frame 0 of every file.  Needs a B200: there is no CPU fallback. */
#include <stdio.h>
#include <stdlib.h>

#include "vt_tracker.h"

#define MAX_STREAMS 16

static int fail(const char* what) {
    fprintf(stderr, "%s: %s\n", what, vt_last_error());
    return 1;
}

int main(int argc, char** argv) {
    if (argc < 10 || argc > 8 + MAX_STREAMS) {
        fprintf(stderr, "usage: %s model.vtw width height x y w h a.nv12 b.nv12 [...]\n", argv[0]);
        return 2;
    }
    const int n = argc - 8;
    vt_config cfg;
    vt_config_default(&cfg);
    cfg.weights_path = argv[1];
    cfg.format = VT_FMT_NV12;
    cfg.width = atoi(argv[2]), cfg.height = atoi(argv[3]);
    cfg.max_targets = n;   /* one target slot per stream */
    cfg.box_overlay = 1;   /* every stream's box is drawn into its own frame */
    cfg.upload_window = 1; /* only the search windows travel over PCIe */
    const size_t frame_bytes = (size_t)cfg.width * (size_t)cfg.height * 3 / 2;
    vt_bbox box;
    box.x = atoi(argv[4]), box.y = atoi(argv[5]), box.width = atoi(argv[6]), box.height = atoi(argv[7]);

    vt_tracker* trk = NULL;
    if (vt_tracker_create(&cfg, &trk) != VT_OK) return fail("vt_tracker_create");
    FILE* f[MAX_STREAMS];
    uint8_t* frames[MAX_STREAMS];
    size_t lens[MAX_STREAMS];
    for (int i = 0; i < n; ++i) {
        void* p = NULL;
        if (vt_alloc_pinned(frame_bytes, &p) != VT_OK) return fail("vt_alloc_pinned");
        frames[i] = (uint8_t*)p, lens[i] = frame_bytes;
        f[i] = fopen(argv[8 + i], "rb");
        if (!f[i]) {
            perror(argv[8 + i]);
            return 1;
        }
    }
    vt_result res[MAX_STREAMS];
    for (long step = 0;; ++step) {
        int got = 0;
        for (int i = 0; i < n; ++i) got += fread(frames[i], 1, frame_bytes, f[i]) == frame_bytes;
        if (got != n) break; /* the shortest file ends the run */
        if (step == 0) {
            for (int i = 0; i < n; ++i)
                if (vt_tracker_init(trk, i, frames[i], frame_bytes, box) != VT_OK) return fail("vt_tracker_init");
            continue;
        }
        if (vt_tracker_update_streams(trk, frames, lens, n, res) != VT_OK) return fail("vt_tracker_update_streams");
        for (int i = 0; i < n; ++i)
            printf("step %ld stream %d: %s score %.4f box (%d, %d, %d, %d)\n", step, i, res[i].success ? "ok  " : "lost", res[i].score, res[i].bbox.x,
                   res[i].bbox.y, res[i].bbox.width, res[i].bbox.height);
    }
    for (int i = 0; i < n; ++i) fclose(f[i]), vt_free_pinned(frames[i]);
    vt_tracker_destroy(trk);
    return 0;
}
