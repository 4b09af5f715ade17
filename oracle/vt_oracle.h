/*
 * vt_oracle.h — CPU ORACLE. TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the reference's per-frame hot path, used only as the checker in
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
 * Nothing in the product (gstreamer_vit_tracker_b200/, include/) may link, import or call it.
 *
 * Pinning:
 *   - pixel code (convert, overlay), TimingStats, SelectionState, TrackerContext follow the
 *     reference source itself (each function cites the file:line it restates) and are checked
 *     against the known-answer vectors derived from that source.
 *   - VitTrack (crop / resize / normalise / net / decode) — PARITY UNPINNED at the reference
 *     boundary: the `vit_tracker` crate is a path dependency that is absent from the reference
 *     tree (Cargo.toml:24, Cargo.lock:1145-1155) and the reference has no tests.  The
 *     restatement follows the published algorithm of OpenCV's TrackerVit (the model name in
 *     src/main.rs:25 is OpenCV's object_tracking_vittrack_2023sep) and is cross-checked
 *     against the executable cv2.TrackerVit 4.13.0 through committed fixtures (tests/golden/).
 */
#ifndef VT_ORACLE_H
#define VT_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- NV12 -> RGB, src/nv12_convert.rs:46-169 ------------------------------------------- */
void vto_nv12_to_rgb(const uint8_t* nv12, size_t len, int width, int height, uint8_t* rgb_out, int threads);
/* YUY2 (packed 4:2:2) -> RGB with the same integer arithmetic; SURVEY.md §8(f) row 1 (parity unpinned: GStreamer videoconvert) */
size_t vto_yuy2_stride(size_t width);
void vto_yuy2_to_rgb(const uint8_t* yuy2, size_t len, size_t width, size_t height, uint8_t* rgb_out, int threads);

/* ---- NV12 overlays (Y plane only), src/nv12_convert.rs:172-343, src/drawing.rs:5-50 ------ */
void vto_draw_rect_nv12(uint8_t* d, int width, int height, int x, int y, int w, int h, int thickness, int brightness);
void vto_draw_crosshair_nv12(uint8_t* d, int width, int height, int cx, int cy, int size, int brightness);
void vto_draw_text_nv12(uint8_t* d, int width, int height, const char* text, int x, int y, int scale, int brightness);
void vto_draw_background_nv12(uint8_t* d, int width, int height, int x, int y, int w, int h, int darkness);
void vto_draw_cursor_nv12(uint8_t* d, int width, int height, int x, int y);
void vto_draw_selection_nv12(uint8_t* d, int width, int height, int start_x, int start_y, int cursor_x, int cursor_y, int selecting_area);
/* returns 0 and fills glyph[7], or -1 for an unknown char (the reference panics, src/drawing.rs:99) */
int vto_get_glyph(int ch, uint8_t glyph[7]);

/* ---- RGB24 overlays, src/drawing_rgb.rs:30-128 ---------------------------------------- */
void vto_draw_background_rgb(uint8_t* d, size_t len, int width, int height, int x, int y, int bw, int bh);
void vto_draw_rect_rgb(uint8_t* d, size_t len, int width, int height, int x, int y, int rw, int rh, int thickness, int r, int g, int b);
void vto_draw_crosshair_rgb(uint8_t* d, size_t len, int width, int height, int cx, int cy, int size, int r, int g, int b);
void vto_draw_cursor_rgb(uint8_t* d, size_t len, int width, int height, int cx, int cy);
void vto_draw_text_rgb(uint8_t* d, size_t len, int width, int height, const char* text, int x, int y, int scale, int luma);
void vto_draw_selection_rgb(uint8_t* d, size_t len, int width, int height, int start_x, int start_y, int cursor_x, int cursor_y, int selecting_area);

/* ---- TimingStats, src/timing_stats.rs:3-60 --------------------------------------------- */
typedef struct vto_timing vto_timing;
vto_timing* vto_timing_new(void);
void vto_timing_free(vto_timing*);
void vto_timing_add_interval(vto_timing*, uint64_t us);
void vto_timing_add_times(vto_timing*, uint64_t conv_us, uint64_t track_us);
double vto_timing_fps(const vto_timing*);
double vto_timing_avg_conv_ms(const vto_timing*);
double vto_timing_avg_track_ms(const vto_timing*);

/* ---- VitTrack (OpenCV TrackerVit semantics) ---------------------- */
typedef struct { int32_t x, y, width, height; } vto_bbox;
typedef struct { int32_t success; float score; vto_bbox bbox; } vto_result;
typedef struct vto_tracker vto_tracker;

/* frame_format: 0 = RGB24 (HWC, channel order as stored), 1 = NV12 (converted internally) */
vto_tracker* vto_tracker_new(const char* weight_path, int threads);
void vto_tracker_free(vto_tracker*);
void vto_tracker_set_threshold(vto_tracker*, float score_threshold);
/* SURVEY.md Appendix A.7 variant switches (all zero = OpenCV 4.13, the default):
 *   pad_plus1      crop padding padR = max(x2-W+1, 0), padB likewise (older OpenCV) instead of max(x2-W, 0)
 *   decode_window  1: cw = 4*floor(sqrt(w*h)) (older) instead of the crop's c = ceil(sqrt(w*h)*4)
 *   window         1: conf * (1 - hann) (older) instead of conf * hann
 *   norm_custom    1: blob = u8*scale[c] + bias[c] instead of (u8/255 - mean_c)/std_c */
typedef struct { int32_t pad_plus1, decode_window, window, norm_custom; float scale[3], bias[3]; } vto_variant;
void vto_tracker_set_variant(vto_tracker*, const vto_variant*);
int vto_tracker_init(vto_tracker*, const uint8_t* rgb, int width, int height, vto_bbox box);
/* returns 0 ok, <0 error (crop lies entirely outside the frame) */
int vto_tracker_update(vto_tracker*, const uint8_t* rgb, int width, int height, vto_result* out);
void vto_tracker_get_rect(const vto_tracker*, vto_bbox* out);
void vto_tracker_set_rect(vto_tracker*, vto_bbox box);
/* diagnostics of the last update: conf*hann[256], size[2*256], offset[2*256], raw conf[256] */
void vto_tracker_last_maps(const vto_tracker*, float* conf_win, float* size_map, float* off_map, float* conf_raw);
/* last search blob (3*256*256 floats, planar CHW) and template blob (3*128*128) */
void vto_tracker_last_blobs(const vto_tracker*, float* search_blob, float* template_blob);

/* building blocks exposed for unit tests */
int vto_crop_square(const uint8_t* rgb, int width, int height, vto_bbox box, int factor, uint8_t* out, int* c_out); /* out may be NULL to query c */
void vto_resize_linear_u8c3(const uint8_t* src, int sw, int sh, uint8_t* dst, int dw, int dh);
void vto_normalize_chw(const uint8_t* hwc, int size, float* chw);
/* net forward on prepared blobs; outputs conf[256] (sigmoid), size[512] (sigmoid), off[512] (raw) */
void vto_net_forward(vto_tracker*, const float* template_blob, const float* search_blob, float* conf, float* size_map, float* off_map);
/* token features after patch-embed + pos and after every block, for layer-wise kernel tests:
   which = 0: embeddings, 1..depth: block outputs, depth+1: final LN.  out[320*D] */
void vto_net_debug_tokens(const vto_tracker*, int which, float* out);
int vto_model_dim(const vto_tracker*, int which); /* 0 D, 1 depth, 2 heads, 3 hidden, 4 head_ch */

/* ---- SelectionState + TrackerContext, src/selection_state.rs, src/tracker_context.rs ---- */
enum { VTO_CMD_MOVE_UP = 0, VTO_CMD_MOVE_DOWN, VTO_CMD_MOVE_LEFT, VTO_CMD_MOVE_RIGHT, VTO_CMD_CONFIRM, VTO_CMD_CANCEL, VTO_CMD_QUIT };
enum { VTO_STATE_SELECT_START = 0, VTO_STATE_SELECT_END, VTO_STATE_TRACKING, VTO_STATE_LOST };
typedef struct vto_context vto_context;
vto_context* vto_context_new(vto_tracker* borrowed_tracker, int width, int height);
void vto_context_free(vto_context*);
void vto_context_handle_command(vto_context*, int cmd, int fast);
/* returns 1 and fills *out when process_frame returns Some(bbox), else 0.  tracker may be NULL for
   pure state-machine tests, in which case `scripted` (success, score, bbox) answers update(). */
int vto_context_process_frame(vto_context*, const uint8_t* rgb, const vto_result* scripted, int scripted_err, vto_bbox* out);
int vto_context_state(const vto_context*);
const char* vto_context_state_name(const vto_context*);
float vto_context_score(const vto_context*);
int vto_context_bbox(const vto_context*, vto_bbox* out);            /* 1 if Some */
void vto_context_selection(const vto_context*, int32_t out[5]);      /* cursor_x, cursor_y, start_x, start_y, phase */
uint64_t vto_context_lost_frames(const vto_context*);

#ifdef __cplusplus
}
#endif
#endif
