/*
 * vt_oracle.c — CPU ORACLE. TEST INFRASTRUCTURE ONLY (see vt_oracle.h for scope and pinning).
 *
 * Plain C restatement of the reference's per-frame hot path.  Every function cites the
 * reference file:line (relative to /root/reference) it follows.  Integer code reproduces the
 * Rust semantics exactly (i32 wrapping add, `as usize` sign extension, saturating_sub,
 * arithmetic >> on negative i32, truncating division).
 *
 * VitTrack part: PARITY UNPINNED at the reference boundary (source absent, no reference
 * tests); restates OpenCV TrackerVit (4.13 behaviour) and is cross-checked against
 * cv2.TrackerVit via tests/golden/.
 */
#define _GNU_SOURCE
#include "vt_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef uint64_t usize; /* Rust usize on the reference's 64-bit target */

static inline usize i32_as_usize(int32_t v) { return (usize)(int64_t)v; }
static inline int32_t wrapping_add_i32(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
static inline usize sat_sub(usize a, usize b) { return a > b ? a - b : 0; }
static inline usize umin(usize a, usize b) { return a < b ? a : b; }

/* ======================================================================================= */
/* NV12 -> RGB                                                                             */
/* ======================================================================================= */

/* src/nv12_convert.rs:41-43 */
static inline uint8_t clamp_u8(int32_t v) { return v < 0 ? 0 : (v > 255 ? 255 : (uint8_t)v); }

/* src/nv12_convert.rs:95-169 (process_row_unsafe); tables of :24-30 are inlined as arithmetic */
static void convert_row(const uint8_t* y_plane, const uint8_t* uv_plane, uint8_t* row_out, size_t row, size_t uv_row, size_t width) {
    const size_t y0 = row * width, uv0 = uv_row * width;
    size_t col = 0;
    while (col + 1 < width) {
        const int32_t u = uv_plane[uv0 + col], v = uv_plane[uv0 + col + 1];
        const int32_t rv = 409 * (v - 128), gu = 100 * (u - 128), gv = 208 * (v - 128), bu = 516 * (u - 128);
        for (int k = 0; k < 2; ++k) {
            const int32_t yv = 298 * ((int32_t)y_plane[y0 + col + k] - 16);
            uint8_t* o = row_out + (col + k) * 3;
            o[0] = clamp_u8((yv + rv + 128) >> 8);
            o[1] = clamp_u8((yv - gu - gv + 128) >> 8);
            o[2] = clamp_u8((yv + bu + 128) >> 8);
        }
        col += 2;
    }
    if (col < width) { /* odd width tail, :150-168 */
        const size_t uvi = uv0 + (col / 2) * 2;
        const int32_t u = uv_plane[uvi], v = uv_plane[uvi + 1];
        const int32_t yv = 298 * ((int32_t)y_plane[y0 + col] - 16);
        uint8_t* o = row_out + col * 3;
        o[0] = clamp_u8((yv + 409 * (v - 128) + 128) >> 8);
        o[1] = clamp_u8((yv - 100 * (u - 128) - 208 * (v - 128) + 128) >> 8);
        o[2] = clamp_u8((yv + 516 * (u - 128) + 128) >> 8);
    }
}

/* src/nv12_convert.rs:46-92 (nv12_full_to_rgb_parallel): short input -> all zeros (:48-50);
 * planes split at w*h, UV row stride == width; parallel over row pairs (:59-88). */
void vto_nv12_to_rgb(const uint8_t* nv12, size_t len, int width, int height, uint8_t* rgb_out, int threads) {
    const size_t w = (size_t)width, h = (size_t)height;
    const size_t ysz = w * h;
    if (len < ysz * 3 / 2) {
        memset(rgb_out, 0, ysz * 3);
        return;
    }
    const uint8_t* yp = nv12;
    const uint8_t* uvp = nv12 + ysz;
    const long pairs = (long)((h + 1) / 2);
    (void)threads;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : 1)
#endif
    for (long p = 0; p < pairs; ++p) {
        const size_t r0 = (size_t)p * 2, r1 = r0 + 1;
        convert_row(yp, uvp, rgb_out + r0 * w * 3, r0, (size_t)p, w);
        if (r1 < h) convert_row(yp, uvp, rgb_out + r1 * w * 3, r1, (size_t)p, w);
    }
}

/* ======================================================================================= */
/* YUY2 -> RGB  (the `videoconvert` step of src/pipeline_ir.rs:27-56)   */
/* ======================================================================================= */
/* The reference delegates this step to GStreamer's videoconvert element, whose source is not in the reference tree: PARITY
 * UNPINNED at that boundary.  Restated with the reference's own BT.601 limited-range integer arithmetic
 * (src/nv12_convert.rs:24-30,124-126,41-43) applied to packed 4:2:2 (bytes Y0 U Y1 V; one chroma pair per two pixels of a row,
 * row stride = width*2 rounded up to 4 bytes as GStreamer lays YUY2 out); a short buffer yields a black frame like :48-50. */
size_t vto_yuy2_stride(size_t width) { return (width * 2 + 3) & ~(size_t)3; }
void vto_yuy2_to_rgb(const uint8_t* yuy2, size_t len, size_t width, size_t height, uint8_t* rgb_out, int threads) {
    const size_t stride = vto_yuy2_stride(width);
    memset(rgb_out, 0, width * height * 3);
    if (len < stride * height) return;
    (void)threads;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : 1)
#endif
    for (long long r = 0; r < (long long)height; ++r) {
        const uint8_t* in = yuy2 + (size_t)r * stride;
        uint8_t* o = rgb_out + (size_t)r * width * 3;
        for (size_t col = 0; col < width; ++col) {
            const uint8_t* q = in + (col >> 1) * 4;
            const int32_t yv = 298 * ((int32_t)q[(col & 1) * 2] - 16), u = q[1], v = q[3];
            o[col * 3 + 0] = clamp_u8((yv + 409 * (v - 128) + 128) >> 8);
            o[col * 3 + 1] = clamp_u8((yv - 100 * (u - 128) - 208 * (v - 128) + 128) >> 8);
            o[col * 3 + 2] = clamp_u8((yv + 516 * (u - 128) + 128) >> 8);
        }
    }
}

/* ======================================================================================= */
/* Glyph table — src/nv12_convert.rs:255-296 == src/drawing.rs:53-94                         */
/* ======================================================================================= */
typedef struct { char ch; uint8_t rows[7]; } glyph_t;
static const glyph_t FONT[40] = {
    {'0', {0x0E, 0x11, 0x13, 0x15, 0x19, 0x11, 0x0E}}, {'1', {0x04, 0x0C, 0x04, 0x04, 0x04, 0x04, 0x0E}},
    {'2', {0x0E, 0x11, 0x01, 0x06, 0x08, 0x10, 0x1F}}, {'3', {0x0E, 0x11, 0x01, 0x06, 0x01, 0x11, 0x0E}},
    {'4', {0x02, 0x06, 0x0A, 0x12, 0x1F, 0x02, 0x02}}, {'5', {0x1F, 0x10, 0x1E, 0x01, 0x01, 0x11, 0x0E}},
    {'6', {0x06, 0x08, 0x10, 0x1E, 0x11, 0x11, 0x0E}}, {'7', {0x1F, 0x01, 0x02, 0x04, 0x08, 0x08, 0x08}},
    {'8', {0x0E, 0x11, 0x11, 0x0E, 0x11, 0x11, 0x0E}}, {'9', {0x0E, 0x11, 0x11, 0x0F, 0x01, 0x02, 0x0C}},
    {'.', {0x00, 0x00, 0x00, 0x00, 0x00, 0x0C, 0x0C}}, {':', {0x00, 0x0C, 0x0C, 0x00, 0x0C, 0x0C, 0x00}},
    {'-', {0x00, 0x00, 0x00, 0x1F, 0x00, 0x00, 0x00}}, {' ', {0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00}},
    {'F', {0x1F, 0x10, 0x1E, 0x10, 0x10, 0x10, 0x10}}, {'P', {0x1E, 0x11, 0x1E, 0x10, 0x10, 0x10, 0x10}},
    {'S', {0x0E, 0x11, 0x10, 0x0E, 0x01, 0x11, 0x0E}}, {'T', {0x1F, 0x04, 0x04, 0x04, 0x04, 0x04, 0x04}},
    {'R', {0x1E, 0x11, 0x1E, 0x14, 0x12, 0x11, 0x11}}, {'A', {0x0E, 0x11, 0x1F, 0x11, 0x11, 0x11, 0x11}},
    {'C', {0x0E, 0x11, 0x10, 0x10, 0x10, 0x11, 0x0E}}, {'K', {0x11, 0x12, 0x14, 0x18, 0x14, 0x12, 0x11}},
    {'I', {0x0E, 0x04, 0x04, 0x04, 0x04, 0x04, 0x0E}}, {'N', {0x11, 0x19, 0x15, 0x13, 0x11, 0x11, 0x11}},
    {'G', {0x0E, 0x11, 0x10, 0x17, 0x11, 0x11, 0x0E}}, {'E', {0x1F, 0x10, 0x1E, 0x10, 0x10, 0x10, 0x1F}},
    {'L', {0x10, 0x10, 0x10, 0x10, 0x10, 0x10, 0x1F}}, {'O', {0x0E, 0x11, 0x11, 0x11, 0x11, 0x11, 0x0E}},
    {'D', {0x1C, 0x12, 0x11, 0x11, 0x11, 0x12, 0x1C}}, {'%', {0x19, 0x1A, 0x04, 0x04, 0x08, 0x0B, 0x13}},
    {'s', {0x00, 0x00, 0x0E, 0x10, 0x0E, 0x01, 0x1E}}, {'c', {0x00, 0x00, 0x0E, 0x10, 0x10, 0x11, 0x0E}},
    {'o', {0x00, 0x00, 0x0E, 0x11, 0x11, 0x11, 0x0E}}, {'r', {0x00, 0x00, 0x16, 0x19, 0x10, 0x10, 0x10}},
    {'e', {0x00, 0x00, 0x0E, 0x11, 0x1F, 0x10, 0x0E}}, {'m', {0x00, 0x00, 0x1A, 0x15, 0x15, 0x11, 0x11}},
    {'t', {0x08, 0x08, 0x1C, 0x08, 0x08, 0x09, 0x06}}, {'k', {0x10, 0x10, 0x12, 0x14, 0x18, 0x14, 0x12}},
    {'n', {0x00, 0x00, 0x16, 0x19, 0x11, 0x11, 0x11}}, {'v', {0x00, 0x00, 0x11, 0x11, 0x11, 0x0A, 0x04}},
};

/* src/drawing.rs:52-100 (get_glyph) — unknown char: the reference panics; here -1 */
int vto_get_glyph(int ch, uint8_t glyph[7]) {
    for (int i = 0; i < 40; ++i)
        if (FONT[i].ch == (char)ch) {
            memcpy(glyph, FONT[i].rows, 7);
            return 0;
        }
    return -1;
}

/* ======================================================================================= */
/* NV12 overlays                                                                           */
/* ======================================================================================= */

/* src/nv12_convert.rs:172-213 */
void vto_draw_rect_nv12(uint8_t* d, int width_, int height_, int x, int y, int w, int h, int thickness_, int brightness) {
    const usize width = (usize)width_, height = (usize)height_, thickness = (usize)thickness_;
    const usize x1 = (usize)(x > 0 ? x : 0), y1 = (usize)(y > 0 ? y : 0);
    const usize x2 = umin(i32_as_usize(wrapping_add_i32(x, w)), sat_sub(width, 1));
    const usize y2 = umin(i32_as_usize(wrapping_add_i32(y, h)), sat_sub(height, 1));
    const uint8_t b = (uint8_t)brightness;
    for (usize t = 0; t < thickness; ++t) {
        if (y1 + t < height)
            for (usize px = x1; px <= x2; ++px) d[(y1 + t) * width + px] = b;
        if (y2 >= t && y2 - t < height)
            for (usize px = x1; px <= x2; ++px) d[(y2 - t) * width + px] = b;
    }
    for (usize py = y1; py <= y2; ++py)
        for (usize t = 0; t < thickness; ++t) {
            if (x1 + t < width) d[py * width + x1 + t] = b;
            if (x2 >= t && x2 - t < width) d[py * width + x2 - t] = b;
        }
}

/* src/nv12_convert.rs:216-242 */
void vto_draw_crosshair_nv12(uint8_t* d, int width_, int height_, int cx_, int cy_, int size_, int brightness) {
    const usize width = (usize)width_, height = (usize)height_;
    const usize cx = (usize)(cx_ > 0 ? cx_ : 0), cy = (usize)(cy_ > 0 ? cy_ : 0), size = i32_as_usize(size_);
    const uint8_t b = (uint8_t)brightness;
    if (cy < height)
        for (usize x = sat_sub(cx, size); x <= umin(cx + size, width - 1); ++x) d[cy * width + x] = b;
    if (cx < width)
        for (usize yy = sat_sub(cy, size); yy <= umin(cy + size, height - 1); ++yy) d[yy * width + cx] = b;
}

/* src/nv12_convert.rs:245-321 — unknown chars draw nothing but still advance (:302,319) */
void vto_draw_text_nv12(uint8_t* d, int width_, int height_, const char* text, int x_, int y_, int scale_, int brightness) {
    const usize width = (usize)width_, height = (usize)height_, y = (usize)y_, scale = (usize)scale_;
    usize cursor_x = (usize)x_;
    const uint8_t b = (uint8_t)brightness;
    for (const char* p = text; *p; ++p) {
        uint8_t g[7];
        if (vto_get_glyph((unsigned char)*p, g) == 0)
            for (usize row = 0; row < 7; ++row)
                for (usize col = 0; col < 5; ++col)
                    if ((g[row] >> (4 - col)) & 1)
                        for (usize dy = 0; dy < scale; ++dy)
                            for (usize dx = 0; dx < scale; ++dx) {
                                const usize px = cursor_x + col * scale + dx, py = y + row * scale + dy;
                                if (px < width && py < height) d[py * width + px] = b;
                            }
        cursor_x += 6 * scale;
    }
}

/* src/nv12_convert.rs:324-343 — u16 multiply, truncating /255 */
void vto_draw_background_nv12(uint8_t* d, int width_, int height_, int x, int y, int w, int h, int darkness) {
    const usize width = (usize)width_, height = (usize)height_;
    const uint16_t factor = (uint16_t)(255 - (uint8_t)darkness);
    for (usize py = (usize)y; py < umin((usize)y + (usize)h, height); ++py)
        for (usize px = (usize)x; px < umin((usize)x + (usize)w, width); ++px) {
            const usize idx = py * width + px;
            d[idx] = (uint8_t)(((uint16_t)d[idx] * factor) / 255);
        }
}

/* src/drawing.rs:5-23 (draw_cursor) */
void vto_draw_cursor_nv12(uint8_t* d, int w_, int h_, int x_, int y_) {
    const usize w = (usize)w_, h = (usize)h_;
    int xc = x_ < 0 ? 0 : (x_ > w_ - 1 ? w_ - 1 : x_);
    int yc = y_ < 0 ? 0 : (y_ > h_ - 1 ? h_ - 1 : y_);
    const usize x = (usize)xc, y = (usize)yc;
    for (usize px = sat_sub(x, 25); px <= umin(x + 25, w - 1); ++px)
        if (!(px >= sat_sub(x, 5) && px <= x + 5)) d[y * w + px] = 255;
    for (usize py = sat_sub(y, 25); py <= umin(y + 25, h - 1); ++py)
        if (!(py >= sat_sub(y, 5) && py <= y + 5)) d[py * w + x] = 255;
}

/* src/drawing.rs:25-50 (draw_selection) */
void vto_draw_selection_nv12(uint8_t* d, int w_, int h_, int start_x, int start_y, int cursor_x, int cursor_y, int selecting_area) {
    if (!selecting_area) return;
    const usize w = (usize)w_, h = (usize)h_;
    int mnx = start_x < cursor_x ? start_x : cursor_x, mny = start_y < cursor_y ? start_y : cursor_y;
    int mxx = start_x > cursor_x ? start_x : cursor_x, mxy = start_y > cursor_y ? start_y : cursor_y;
    const usize x1 = (usize)(mnx > 0 ? mnx : 0), y1 = (usize)(mny > 0 ? mny : 0);
    const usize x2 = umin(i32_as_usize(mxx), w - 1), y2 = umin(i32_as_usize(mxy), h - 1);
    for (usize x = x1; x <= x2; ++x)
        if ((x / 6) % 2 == 0) {
            d[y1 * w + x] = 255;
            d[y2 * w + x] = 255;
        }
    for (usize y = y1; y <= y2; ++y)
        if ((y / 6) % 2 == 0) {
            d[y * w + x1] = 255;
            d[y * w + x2] = 255;
        }
}

/* ======================================================================================= */
/* RGB24 overlays — src/drawing_rgb.rs                                                      */
/* ======================================================================================= */

/* :18-28 (set_pixel_rgb_color); :5-15 (set_pixel_rgb) is the r=g=b case */
static inline void set_px(uint8_t* d, size_t len, int w, int h, int x, int y, uint8_t r, uint8_t g, uint8_t b) {
    if (x < 0 || y < 0 || x >= w || y >= h) return;
    const size_t off = ((size_t)y * (size_t)w + (size_t)x) * 3;
    if (off + 2 < len) {
        d[off] = r;
        d[off + 1] = g;
        d[off + 2] = b;
    }
}

/* :30-53 — memset-style fill with 30 */
void vto_draw_background_rgb(uint8_t* d, size_t len, int w_, int h_, int x, int y, int bw, int bh) {
    const usize w = (usize)w_, h = (usize)h_;
    const usize xs = (usize)(x > 0 ? x : 0), xe = umin(i32_as_usize(wrapping_add_i32(x, bw)), w);
    const usize ys = (usize)(y > 0 ? y : 0), ye = umin(i32_as_usize(wrapping_add_i32(y, bh)), h);
    if (xe <= xs) return; /* the reference would underflow/panic here */
    const usize row_bytes = (xe - xs) * 3;
    for (usize row = ys; row < ye; ++row) {
        const usize off = (row * w + xs) * 3;
        if (off + row_bytes <= len) memset(d + off, 30, row_bytes);
    }
}

/* :55-66 — exclusive geometry x..x+rw-1 */
void vto_draw_rect_rgb(uint8_t* d, size_t len, int w, int h, int x, int y, int rw, int rh, int thickness, int r, int g, int b) {
    for (int t = 0; t < thickness; ++t) {
        for (int i = 0; i < rw; ++i) {
            set_px(d, len, w, h, x + i, y + t, (uint8_t)r, (uint8_t)g, (uint8_t)b);
            set_px(d, len, w, h, x + i, y + rh - 1 - t, (uint8_t)r, (uint8_t)g, (uint8_t)b);
        }
        for (int i = 0; i < rh; ++i) {
            set_px(d, len, w, h, x + t, y + i, (uint8_t)r, (uint8_t)g, (uint8_t)b);
            set_px(d, len, w, h, x + rw - 1 - t, y + i, (uint8_t)r, (uint8_t)g, (uint8_t)b);
        }
    }
}

/* :68-73 */
void vto_draw_crosshair_rgb(uint8_t* d, size_t len, int w, int h, int cx, int cy, int size, int r, int g, int b) {
    for (int i = -size; i <= size; ++i) {
        set_px(d, len, w, h, cx + i, cy, (uint8_t)r, (uint8_t)g, (uint8_t)b);
        set_px(d, len, w, h, cx, cy + i, (uint8_t)r, (uint8_t)g, (uint8_t)b);
    }
}

/* :75-84 — gap 5..=25, green */
void vto_draw_cursor_rgb(uint8_t* d, size_t len, int w, int h, int cx, int cy) {
    for (int i = 5; i <= 25; ++i) {
        set_px(d, len, w, h, cx + i, cy, 0, 255, 0);
        set_px(d, len, w, h, cx - i, cy, 0, 255, 0);
        set_px(d, len, w, h, cx, cy + i, 0, 255, 0);
        set_px(d, len, w, h, cx, cy - i, 0, 255, 0);
    }
}

/* :86-104 — get_glyph panics on unknown chars (src/drawing.rs:99); here they are skipped */
void vto_draw_text_rgb(uint8_t* d, size_t len, int w, int h, const char* text, int x, int y, int scale, int luma) {
    int cx = x;
    for (const char* p = text; *p; ++p) {
        uint8_t g[7];
        if (vto_get_glyph((unsigned char)*p, g) == 0)
            for (int gy = 0; gy < 7; ++gy)
                for (int gx = 0; gx < 5; ++gx)
                    if ((g[gy] >> (4 - gx)) & 1)
                        for (int sy = 0; sy < scale; ++sy)
                            for (int sx = 0; sx < scale; ++sx)
                                set_px(d, len, w, h, cx + gx * scale + sx, y + gy * scale + sy, (uint8_t)luma, (uint8_t)luma, (uint8_t)luma);
        cx += 6 * scale;
    }
}

/* :106-128 — yellow dashed */
void vto_draw_selection_rgb(uint8_t* d, size_t len, int w, int h, int start_x, int start_y, int cursor_x, int cursor_y, int selecting_area) {
    if (!selecting_area) return;
    int x1 = start_x < cursor_x ? start_x : cursor_x, y1 = start_y < cursor_y ? start_y : cursor_y;
    int x2 = start_x > cursor_x ? start_x : cursor_x, y2 = start_y > cursor_y ? start_y : cursor_y;
    if (x1 < 0) x1 = 0;
    if (y1 < 0) y1 = 0;
    if (x2 > w - 1) x2 = w - 1;
    if (y2 > h - 1) y2 = h - 1;
    for (int x = x1; x <= x2; ++x)
        if ((x / 6) % 2 == 0) {
            set_px(d, len, w, h, x, y1, 255, 255, 0);
            set_px(d, len, w, h, x, y2, 255, 255, 0);
        }
    for (int y = y1; y <= y2; ++y)
        if ((y / 6) % 2 == 0) {
            set_px(d, len, w, h, x1, y, 255, 255, 0);
            set_px(d, len, w, h, x2, y, 255, 255, 0);
        }
}

/* ======================================================================================= */
/* TimingStats — src/timing_stats.rs:3-60                                                   */
/* ======================================================================================= */
#define VTO_WIN 120
typedef struct { uint64_t v[VTO_WIN]; int head, len; } ring_t;
struct vto_timing { ring_t intervals, conv, track; };

static void ring_push(ring_t* r, uint64_t x) { /* pop_front when len >= 120, then push_back */
    if (r->len >= VTO_WIN) {
        r->head = (r->head + 1) % VTO_WIN;
        r->len--;
    }
    r->v[(r->head + r->len) % VTO_WIN] = x;
    r->len++;
}
static double ring_mean(const ring_t* r) {
    uint64_t s = 0;
    for (int i = 0; i < r->len; ++i) s += r->v[(r->head + i) % VTO_WIN];
    return (double)s / (double)r->len;
}
vto_timing* vto_timing_new(void) { return (vto_timing*)calloc(1, sizeof(vto_timing)); }
void vto_timing_free(vto_timing* t) { free(t); }
void vto_timing_add_interval(vto_timing* t, uint64_t us) { ring_push(&t->intervals, us); }
void vto_timing_add_times(vto_timing* t, uint64_t c, uint64_t k) {
    ring_push(&t->conv, c);
    ring_push(&t->track, k);
}
double vto_timing_fps(const vto_timing* t) { /* :36-46 */
    if (t->intervals.len == 0) return 0.0;
    const double avg = ring_mean(&t->intervals);
    return avg > 0.0 ? 1000000.0 / avg : 0.0;
}
double vto_timing_avg_conv_ms(const vto_timing* t) { return t->conv.len ? ring_mean(&t->conv) / 1000.0 : 0.0; }
double vto_timing_avg_track_ms(const vto_timing* t) { return t->track.len ? ring_mean(&t->track) / 1000.0 : 0.0; }

/* ======================================================================================= */
/* VitTrack — OpenCV TrackerVit semantics; PARITY UNPINNED upstream   */
/* ======================================================================================= */
#define NTZ 64
#define NTX 256
#define NTOK 320
#define PATCH_K 768

typedef struct {
    const float *ln1_g, *ln1_b, *qkv_w, *qkv_b, *proj_w, *proj_b, *ln2_g, *ln2_b, *fc1_w, *fc1_b, *fc2_w, *fc2_b;
} blk_t;

struct vto_tracker {
    int D, depth, heads, hidden, head_ch, threads;
    float* storage;
    const float *patch_w, *patch_b, *pos_z, *pos_x, *lnf_g, *lnf_b, *h1_w, *h1_b, *h2_w, *h2_b;
    blk_t* blk;
    float threshold;
    vto_variant var; /* App. A.7 switches */
    vto_bbox rect_last;
    float norm_lut[3][256];
    float hann[256];
    float* template_blob; /* 3*128*128 */
    float* search_blob;   /* 3*256*256 */
    float conf_raw[256], conf_win[256], size_map[512], off_map[512];
    float* dbg; /* (depth+2) * 320 * D */
};

static const double MEANV[3] = {0.485, 0.456, 0.406};
static const double STDV[3] = {0.229, 0.224, 0.225};

vto_tracker* vto_tracker_new(const char* path, int threads) {
    FILE* f = fopen(path, "rb");
    if (!f) return NULL;
    char magic[4];
    int32_t hdr[7];
    if (fread(magic, 1, 4, f) != 4 || memcmp(magic, "VTW1", 4) != 0 || fread(hdr, 4, 7, f) != 7) {
        fclose(f);
        return NULL;
    }
    /* same shape limits as the library's loader: a corrupt header must not become a huge allocation */
    if (hdr[0] <= 0 || hdr[0] > 1024 || hdr[1] <= 0 || hdr[1] > 64 || hdr[2] <= 0 || hdr[0] % hdr[2] || hdr[3] <= 0 || hdr[3] > 8192 ||
        hdr[4] <= 0 || hdr[4] > 1024) {
        fclose(f);
        return NULL;
    }
    vto_tracker* t = (vto_tracker*)calloc(1, sizeof(*t));
    t->D = hdr[0], t->depth = hdr[1], t->heads = hdr[2], t->hidden = hdr[3], t->head_ch = hdr[4];
    t->threads = threads > 0 ? threads : 1;
    const size_t D = t->D, H = t->hidden, C = t->head_ch;
    size_t n = D * PATCH_K + D + NTZ * D + NTX * D + (size_t)t->depth * (4 * D + 3 * D * D + 3 * D + D * D + D + H * D + H + D * H + D) + 2 * D + C * D * 9 + C + 5 * C + 5;
    t->storage = (float*)malloc(n * sizeof(float));
    if (fread(t->storage, sizeof(float), n, f) != n) {
        fclose(f);
        free(t->storage);
        free(t);
        return NULL;
    }
    fclose(f);
    const float* p = t->storage;
#define TAKE(dst, cnt) do { dst = p; p += (cnt); } while (0)
    TAKE(t->patch_w, D * PATCH_K);
    TAKE(t->patch_b, D);
    TAKE(t->pos_z, NTZ * D);
    TAKE(t->pos_x, NTX * D);
    t->blk = (blk_t*)calloc((size_t)t->depth, sizeof(blk_t));
    for (int i = 0; i < t->depth; ++i) {
        blk_t* b = &t->blk[i];
        TAKE(b->ln1_g, D); TAKE(b->ln1_b, D);
        TAKE(b->qkv_w, 3 * D * D); TAKE(b->qkv_b, 3 * D);
        TAKE(b->proj_w, D * D); TAKE(b->proj_b, D);
        TAKE(b->ln2_g, D); TAKE(b->ln2_b, D);
        TAKE(b->fc1_w, H * D); TAKE(b->fc1_b, H);
        TAKE(b->fc2_w, D * H); TAKE(b->fc2_b, D);
    }
    TAKE(t->lnf_g, D); TAKE(t->lnf_b, D);
    TAKE(t->h1_w, C * D * 9); TAKE(t->h1_b, C);
    TAKE(t->h2_w, 5 * C); TAKE(t->h2_b, 5);
#undef TAKE
    t->threshold = 0.20f; /* cv2 default tracking_score_threshold */
    /* A.4: blob = (u8/255 - mean_c)/std_c, evaluated in double and rounded once to fp32 */
    for (int c = 0; c < 3; ++c)
        for (int v = 0; v < 256; ++v) t->norm_lut[c][v] = (float)(((double)v / 255.0 - MEANV[c]) / STDV[c]);
    /* A.5: hann1d[i] = 0.5*(1 - cos(2*pi*(i+1)/17)), fp32 as OpenCV computes it */
    float h1[16];
    for (int i = 0; i < 16; ++i) h1[i] = 0.5f * (1.f - cosf((float)(2 * M_PI / 17) * (float)(i + 1)));
    for (int y = 0; y < 16; ++y)
        for (int x = 0; x < 16; ++x) t->hann[y * 16 + x] = h1[y] * h1[x];
    t->template_blob = (float*)calloc(3 * 128 * 128, sizeof(float));
    t->search_blob = (float*)calloc(3 * 256 * 256, sizeof(float));
    t->dbg = (float*)calloc((size_t)(t->depth + 2) * NTOK * D, sizeof(float));
    return t;
}

void vto_tracker_free(vto_tracker* t) {
    if (!t) return;
    free(t->storage);
    free(t->blk);
    free(t->template_blob);
    free(t->search_blob);
    free(t->dbg);
    free(t);
}
void vto_tracker_set_threshold(vto_tracker* t, float s) { t->threshold = s; }
/* App. A.7: rebuilds the normalisation LUT and the window table for the chosen variant */
void vto_tracker_set_variant(vto_tracker* t, const vto_variant* v) {
    t->var = *v;
    for (int c = 0; c < 3; ++c)
        for (int i = 0; i < 256; ++i)
            t->norm_lut[c][i] = v->norm_custom ? (float)((double)i * (double)v->scale[c] + (double)v->bias[c])
                                               : (float)(((double)i / 255.0 - MEANV[c]) / STDV[c]);
    float h1[16];
    for (int i = 0; i < 16; ++i) h1[i] = 0.5f * (1.f - cosf((float)(2 * M_PI / 17) * (float)(i + 1)));
    for (int y = 0; y < 16; ++y)
        for (int x = 0; x < 16; ++x) t->hann[y * 16 + x] = v->window ? 1.f - h1[y] * h1[x] : h1[y] * h1[x];
}
void vto_tracker_get_rect(const vto_tracker* t, vto_bbox* o) { *o = t->rect_last; }
void vto_tracker_set_rect(vto_tracker* t, vto_bbox b) { t->rect_last = b; }
int vto_model_dim(const vto_tracker* t, int which) {
    switch (which) {
        case 0: return t->D;
        case 1: return t->depth;
        case 2: return t->heads;
        case 3: return t->hidden;
        default: return t->head_ch;
    }
}

/* A.1: c = ceil(sqrt(w*h)*factor) */
static int crop_size(vto_bbox b, int factor) { return (int)ceil(sqrt((double)(b.width * b.height)) * (double)factor); }

/* A.1 crop with zero border.  Returns 0, or -1 when the crop lies entirely outside the frame
 * (cv2 throws an ROI assertion there).  out: c*c*3 */
static int crop_square_v(const uint8_t* rgb, int W, int H, vto_bbox box, int factor, uint8_t* out, int* c_out, int pad_plus1);
int vto_crop_square(const uint8_t* rgb, int W, int H, vto_bbox box, int factor, uint8_t* out, int* c_out) {
    return crop_square_v(rgb, W, H, box, factor, out, c_out, 0);
}
/* pad_plus1 (App. A.7, older OpenCV): x2_pad = max(x2 - W + 1, 0), y2_pad likewise — a crop that reaches the right / bottom edge
 * treats the last column / row of the frame as border */
static int crop_square_v(const uint8_t* rgb, int W, int H, vto_bbox box, int factor, uint8_t* out, int* c_out, int pad_plus1) {
    const int c = crop_size(box, factor);
    if (c_out) *c_out = c;
    if (c <= 0) return -1;
    const int x1 = box.x + (box.width - c) / 2, y1 = box.y + (box.height - c) / 2; /* C truncating division */
    const int x2 = x1 + c, y2 = y1 + c;
    const int pl = x1 < 0 ? -x1 : 0, pt = y1 < 0 ? -y1 : 0;
    const int pr = x2 - W + pad_plus1 > 0 ? x2 - W + pad_plus1 : 0, pb = y2 - H + pad_plus1 > 0 ? y2 - H + pad_plus1 : 0;
    const int rw = c - pl - pr, rh = c - pt - pb;
    if (rw <= 0 || rh <= 0) return -1;
    if (!out) return 0;
    memset(out, 0, (size_t)c * c * 3);
    for (int r = 0; r < rh; ++r)
        memcpy(out + ((size_t)(r + pt) * c + pl) * 3, rgb + ((size_t)(y1 + pt + r) * W + (x1 + pl)) * 3, (size_t)rw * 3);
    return 0;
}

/* A.3: OpenCV INTER_LINEAR on 8UC3 (fixed point, 11-bit coefficients).  Horizontal taps clamp the
 * fractional part at the borders; vertical taps keep the coefficients and clamp the ROW INDEX
 * (that asymmetry is what makes up-scales bit-exact against cv2.resize, verified for
 * src 1..1000 -> 128/256). */
static void lin_coef(int s, int d, int clamp_frac, int* ofs, int16_t* a0, int16_t* a1) {
    const double scale = 1.0 / ((double)d / (double)s);
    for (int i = 0; i < d; ++i) {
        float f = (float)(((double)i + 0.5) * scale - 0.5);
        int ix = (int)floorf(f);
        f -= (float)ix;
        if (clamp_frac) {
            if (ix < 0) ix = 0, f = 0.f;
            if (ix >= s - 1) ix = s - 1, f = 0.f;
        }
        ofs[i] = ix;
        a0[i] = (int16_t)lrintf((1.f - f) * 2048.f);
        a1[i] = (int16_t)lrintf(f * 2048.f);
    }
}
void vto_resize_linear_u8c3(const uint8_t* src, int sw, int sh, uint8_t* dst, int dw, int dh) {
    int* xo = (int*)malloc(sizeof(int) * (size_t)(dw + dh));
    int* yo = xo + dw;
    int16_t* ca = (int16_t*)malloc(sizeof(int16_t) * 2 * (size_t)(dw + dh));
    int16_t *a0 = ca, *a1 = ca + dw, *b0 = ca + 2 * dw, *b1 = ca + 2 * dw + dh;
    lin_coef(sw, dw, 1, xo, a0, a1);
    lin_coef(sh, dh, 0, yo, b0, b1);
    for (int dy = 0; dy < dh; ++dy) {
        int r0 = yo[dy], r1 = yo[dy] + 1;
        r0 = r0 < 0 ? 0 : (r0 > sh - 1 ? sh - 1 : r0);
        r1 = r1 < 0 ? 0 : (r1 > sh - 1 ? sh - 1 : r1);
        const uint8_t *s0 = src + (size_t)r0 * sw * 3, *s1 = src + (size_t)r1 * sw * 3;
        for (int dx = 0; dx < dw; ++dx) {
            const int x0 = xo[dx], x1 = x0 + 1 < sw ? x0 + 1 : sw - 1;
            for (int ch = 0; ch < 3; ++ch) {
                const int t0 = s0[x0 * 3 + ch] * a0[dx] + s0[x1 * 3 + ch] * a1[dx];
                const int t1 = s1[x0 * 3 + ch] * a0[dx] + s1[x1 * 3 + ch] * a1[dx];
                dst[((size_t)dy * dw + dx) * 3 + ch] = (uint8_t)((((b0[dy] * (t0 >> 4)) >> 16) + ((b1[dy] * (t1 >> 4)) >> 16) + 2) >> 2);
            }
        }
    }
    free(xo);
    free(ca);
}

static void normalize_with_lut(const float lut[3][256], const uint8_t* hwc, int size, float* chw) {
    const size_t n = (size_t)size * size;
    for (size_t i = 0; i < n; ++i)
        for (int c = 0; c < 3; ++c) chw[(size_t)c * n + i] = lut[c][hwc[i * 3 + c]];
}
/* A.4 — channels in memory order, no swap */
void vto_normalize_chw(const uint8_t* hwc, int size, float* chw) {
    float lut[3][256];
    for (int c = 0; c < 3; ++c)
        for (int v = 0; v < 256; ++v) lut[c][v] = (float)(((double)v / 255.0 - MEANV[c]) / STDV[c]);
    normalize_with_lut(lut, hwc, size, chw);
}

/* ---- fp32 network ---------------------------------------------------------------------- */
static void gemm_nt(const float* A, const float* W, const float* bias, float* C, int M, int N, int K, int threads) {
    (void)threads;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(threads)
#endif
    for (int m = 0; m < M; ++m) {
        const float* a = A + (size_t)m * K;
        for (int n = 0; n < N; ++n) {
            const float* w = W + (size_t)n * K;
            float acc = 0.f;
#pragma omp simd reduction(+ : acc)
            for (int k = 0; k < K; ++k) acc += a[k] * w[k];
            C[(size_t)m * N + n] = acc + (bias ? bias[n] : 0.f);
        }
    }
}

static void layernorm(const float* x, const float* g, const float* b, float* y, int M, int D) {
    for (int m = 0; m < M; ++m) {
        const float* r = x + (size_t)m * D;
        float mean = 0.f;
        for (int i = 0; i < D; ++i) mean += r[i];
        mean /= (float)D;
        float var = 0.f;
        for (int i = 0; i < D; ++i) var += (r[i] - mean) * (r[i] - mean);
        var /= (float)D;
        const float inv = 1.f / sqrtf(var + 1e-6f);
        for (int i = 0; i < D; ++i) y[(size_t)m * D + i] = (r[i] - mean) * inv * g[i] + b[i];
    }
}

static inline float gelu_exact(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }
static inline float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

static void patch_tokens(const float* blob, int size, float* A) { /* A[(ty*nt+tx), c*256+py*16+px] */
    const int nt = size / 16;
    for (int ty = 0; ty < nt; ++ty)
        for (int tx = 0; tx < nt; ++tx)
            for (int c = 0; c < 3; ++c)
                for (int py = 0; py < 16; ++py)
                    memcpy(A + (size_t)(ty * nt + tx) * PATCH_K + c * 256 + py * 16,
                           blob + (size_t)c * size * size + (size_t)(ty * 16 + py) * size + tx * 16, 16 * sizeof(float));
}

void vto_net_forward(vto_tracker* t, const float* zblob, const float* xblob, float* conf, float* size_map, float* off_map) {
    const int D = t->D, H = t->hidden, C = t->head_ch, heads = t->heads, dh = D / heads, th = t->threads;
    float* A = (float*)malloc(sizeof(float) * NTOK * PATCH_K);
    float* x = (float*)malloc(sizeof(float) * NTOK * D);
    float* y = (float*)malloc(sizeof(float) * NTOK * D);
    float* qkv = (float*)malloc(sizeof(float) * NTOK * 3 * D);
    float* att = (float*)malloc(sizeof(float) * NTOK * D);
    float* hid = (float*)malloc(sizeof(float) * NTOK * H);
    float* tmp = (float*)malloc(sizeof(float) * NTOK * D);
    patch_tokens(zblob, 128, A);
    patch_tokens(xblob, 256, A + (size_t)NTZ * PATCH_K);
    gemm_nt(A, t->patch_w, t->patch_b, x, NTOK, D, PATCH_K, th);
    for (int m = 0; m < NTOK; ++m) {
        const float* pos = m < NTZ ? t->pos_z + (size_t)m * D : t->pos_x + (size_t)(m - NTZ) * D;
        for (int i = 0; i < D; ++i) x[(size_t)m * D + i] += pos[i];
    }
    memcpy(t->dbg, x, sizeof(float) * NTOK * D);
    const float scale = 1.f / sqrtf((float)dh);
    for (int l = 0; l < t->depth; ++l) {
        const blk_t* b = &t->blk[l];
        layernorm(x, b->ln1_g, b->ln1_b, y, NTOK, D);
        gemm_nt(y, b->qkv_w, b->qkv_b, qkv, NTOK, 3 * D, D, th);
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(th) collapse(2)
#endif
        for (int h = 0; h < heads; ++h)
            for (int i = 0; i < NTOK; ++i) {
                float s[NTOK];
                const float* q = qkv + (size_t)i * 3 * D + h * dh;
                float mx = -INFINITY;
                for (int j = 0; j < NTOK; ++j) {
                    const float* k = qkv + (size_t)j * 3 * D + D + h * dh;
                    float acc = 0.f;
                    for (int d = 0; d < dh; ++d) acc += q[d] * k[d];
                    s[j] = acc * scale;
                    mx = s[j] > mx ? s[j] : mx;
                }
                float sum = 0.f;
                for (int j = 0; j < NTOK; ++j) {
                    s[j] = expf(s[j] - mx);
                    sum += s[j];
                }
                const float inv = 1.f / sum;
                float* o = att + (size_t)i * D + h * dh;
                for (int d = 0; d < dh; ++d) o[d] = 0.f;
                for (int j = 0; j < NTOK; ++j) {
                    const float p = s[j] * inv;
                    const float* v = qkv + (size_t)j * 3 * D + 2 * D + h * dh;
                    for (int d = 0; d < dh; ++d) o[d] += p * v[d];
                }
            }
        gemm_nt(att, b->proj_w, b->proj_b, tmp, NTOK, D, D, th);
        for (size_t i = 0; i < (size_t)NTOK * D; ++i) x[i] += tmp[i];
        layernorm(x, b->ln2_g, b->ln2_b, y, NTOK, D);
        gemm_nt(y, b->fc1_w, b->fc1_b, hid, NTOK, H, D, th);
        for (size_t i = 0; i < (size_t)NTOK * H; ++i) hid[i] = gelu_exact(hid[i]);
        gemm_nt(hid, b->fc2_w, b->fc2_b, tmp, NTOK, D, H, th);
        for (size_t i = 0; i < (size_t)NTOK * D; ++i) x[i] += tmp[i];
        memcpy(t->dbg + (size_t)(l + 1) * NTOK * D, x, sizeof(float) * NTOK * D);
    }
    layernorm(x, t->lnf_g, t->lnf_b, y, NTOK, D);
    memcpy(t->dbg + (size_t)(t->depth + 1) * NTOK * D, y, sizeof(float) * NTOK * D);
    /* head: search tokens -> [D,16,16]; conv3x3 (pad 1) + ReLU; conv1x1 -> 5 maps */
    const float* f = y + (size_t)NTZ * D; /* token (yy*16+xx) feature d */
    float* h1 = (float*)malloc(sizeof(float) * NTX * C);
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(th)
#endif
    for (int p = 0; p < NTX; ++p) {
        const int py = p / 16, px = p % 16;
        for (int c = 0; c < C; ++c) {
            float acc = t->h1_b[c];
            for (int ky = 0; ky < 3; ++ky) {
                const int yy = py + ky - 1;
                if (yy < 0 || yy > 15) continue;
                for (int kx = 0; kx < 3; ++kx) {
                    const int xx = px + kx - 1;
                    if (xx < 0 || xx > 15) continue;
                    const float* fv = f + (size_t)(yy * 16 + xx) * D;
                    const float* wv = t->h1_w + (size_t)c * D * 9 + ky * 3 + kx; /* [C, D, 3, 3] */
                    float a2 = 0.f;
                    for (int d = 0; d < D; ++d) a2 += fv[d] * wv[(size_t)d * 9];
                    acc += a2;
                }
            }
            h1[(size_t)p * C + c] = acc > 0.f ? acc : 0.f;
        }
    }
    for (int p = 0; p < NTX; ++p) {
        float o[5];
        for (int k = 0; k < 5; ++k) {
            float acc = t->h2_b[k];
            for (int c = 0; c < C; ++c) acc += h1[(size_t)p * C + c] * t->h2_w[k * C + c];
            o[k] = acc;
        }
        conf[p] = sigmoidf_(o[0]);
        size_map[p] = sigmoidf_(o[1]);
        size_map[256 + p] = sigmoidf_(o[2]);
        off_map[p] = o[3];
        off_map[256 + p] = o[4];
    }
    free(h1); free(A); free(x); free(y); free(qkv); free(att); free(hid); free(tmp);
}

void vto_net_debug_tokens(const vto_tracker* t, int which, float* out) {
    memcpy(out, t->dbg + (size_t)which * NTOK * t->D, sizeof(float) * NTOK * t->D);
}

static int make_blob(vto_tracker* t, const uint8_t* rgb, int W, int H, vto_bbox box, int factor, int size, float* blob) {
    int c;
    if (crop_square_v(rgb, W, H, box, factor, NULL, &c, t->var.pad_plus1) != 0) return -1;
    uint8_t* crop = (uint8_t*)malloc((size_t)c * c * 3);
    uint8_t* rs = (uint8_t*)malloc((size_t)size * size * 3);
    crop_square_v(rgb, W, H, box, factor, crop, NULL, t->var.pad_plus1);
    vto_resize_linear_u8c3(crop, c, c, rs, size, size);
    normalize_with_lut(t->norm_lut, rs, size, blob);
    free(crop);
    free(rs);
    return 0;
}

/* A.2 */
int vto_tracker_init(vto_tracker* t, const uint8_t* rgb, int W, int H, vto_bbox box) {
    if (make_blob(t, rgb, W, H, box, 2, 128, t->template_blob) != 0) return -1;
    t->rect_last = box;
    return 0;
}

/* A.5 / A.6 */
int vto_tracker_update(vto_tracker* t, const uint8_t* rgb, int W, int H, vto_result* out) {
    out->success = 0;
    out->score = 0.f;
    out->bbox.x = out->bbox.y = out->bbox.width = out->bbox.height = 0;
    if (make_blob(t, rgb, W, H, t->rect_last, 4, 256, t->search_blob) != 0) return -1;
    vto_net_forward(t, t->template_blob, t->search_blob, t->conf_raw, t->size_map, t->off_map);
    int best = 0;
    float bv = -INFINITY;
    for (int i = 0; i < 256; ++i) {
        t->conf_win[i] = t->conf_raw[i] * t->hann[i];
        if (t->conf_win[i] > bv) bv = t->conf_win[i], best = i; /* first row-major maximum */
    }
    out->score = bv;
    if (bv >= t->threshold) {
        const int my = best / 16, mx = best % 16;
        const float cx = ((float)mx + t->off_map[best]) / 16.f;
        const float cy = ((float)my + t->off_map[256 + best]) / 16.f;
        const float bw = t->size_map[best], bh = t->size_map[256 + best];
        const vto_bbox L = t->rect_last;
        /* A.6: the crop's own c; App. A.7 decode_window = 1: the older 4 * floor(sqrt(w*h)) */
        const int cw = t->var.decode_window ? 4 * (int)floor(sqrt((double)(L.width * L.height))) : crop_size(L, 4);
        const int x0 = L.x + (L.width - cw) / 2, y0 = L.y + (L.height - cw) / 2;
        vto_bbox r;
        r.x = (int)floorf((cx - bw / 2.f) * (float)cw + (float)x0);
        r.y = (int)floorf((cy - bh / 2.f) * (float)cw + (float)y0);
        r.width = (int)floorf(bw * (float)cw);
        r.height = (int)floorf(bh * (float)cw);
        t->rect_last = r;
        out->bbox = r;
        out->success = 1;
    }
    return 0;
}

void vto_tracker_last_maps(const vto_tracker* t, float* cw, float* sm, float* om, float* cr) {
    if (cw) memcpy(cw, t->conf_win, sizeof(t->conf_win));
    if (sm) memcpy(sm, t->size_map, sizeof(t->size_map));
    if (om) memcpy(om, t->off_map, sizeof(t->off_map));
    if (cr) memcpy(cr, t->conf_raw, sizeof(t->conf_raw));
}
void vto_tracker_last_blobs(const vto_tracker* t, float* sb, float* tb) {
    if (sb) memcpy(sb, t->search_blob, sizeof(float) * 3 * 256 * 256);
    if (tb) memcpy(tb, t->template_blob, sizeof(float) * 3 * 128 * 128);
}

/* ======================================================================================= */
/* SelectionState (src/selection_state.rs:1-45) + TrackerContext (src/tracker_context.rs)   */
/* ======================================================================================= */
typedef struct { int32_t cursor_x, cursor_y, start_x, start_y, phase, step, fast_step; } sel_t;
struct vto_context {
    vto_tracker* tracker;
    int state; /* 0 Selecting, 1 Tracking, 2 Lost */
    uint64_t lost_frames;
    sel_t sel;
    int has_bbox;
    vto_bbox bbox;
    float score;
    int32_t fw, fh;
    int pending_confirm;
};

static void sel_new(sel_t* s, int w, int h) { /* selection_state.rs:21-31 */
    s->cursor_x = s->start_x = w / 2;
    s->cursor_y = s->start_y = h / 2;
    s->phase = 0;
    s->step = 10;
    s->fast_step = 50;
}
static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static void sel_move(sel_t* s, int dx, int dy, int fast, int w, int h) { /* :33-37 */
    const int step = fast ? s->fast_step : s->step;
    s->cursor_x = clampi(s->cursor_x + dx * step, 0, w - 1);
    s->cursor_y = clampi(s->cursor_y + dy * step, 0, h - 1);
}
static vto_bbox sel_bbox(const sel_t* s) { /* :39-45, min side 20 */
    vto_bbox b;
    b.x = s->start_x < s->cursor_x ? s->start_x : s->cursor_x;
    b.y = s->start_y < s->cursor_y ? s->start_y : s->cursor_y;
    const int w = abs(s->start_x - s->cursor_x), h = abs(s->start_y - s->cursor_y);
    b.width = w > 20 ? w : 20;
    b.height = h > 20 ? h : 20;
    return b;
}

vto_context* vto_context_new(vto_tracker* tr, int w, int h) { /* tracker_context.rs:19-34 */
    vto_context* c = (vto_context*)calloc(1, sizeof(*c));
    c->tracker = tr;
    c->fw = w;
    c->fh = h;
    sel_new(&c->sel, w, h);
    return c;
}
void vto_context_free(vto_context* c) { free(c); }

void vto_context_handle_command(vto_context* c, int cmd, int fast) { /* :36-61 */
    switch (cmd) {
        case VTO_CMD_MOVE_UP: sel_move(&c->sel, 0, -1, fast, c->fw, c->fh); break;
        case VTO_CMD_MOVE_DOWN: sel_move(&c->sel, 0, 1, fast, c->fw, c->fh); break;
        case VTO_CMD_MOVE_LEFT: sel_move(&c->sel, -1, 0, fast, c->fw, c->fh); break;
        case VTO_CMD_MOVE_RIGHT: sel_move(&c->sel, 1, 0, fast, c->fw, c->fh); break;
        case VTO_CMD_CONFIRM: c->pending_confirm = 1; break;
        case VTO_CMD_CANCEL:
            c->state = 0;
            sel_new(&c->sel, c->fw, c->fh);
            c->has_bbox = 0;
            break;
        default: break;
    }
}

static int ctx_update(vto_context* c, const uint8_t* rgb, const vto_result* scripted, int scripted_err, vto_result* r) {
    if (c->tracker) return vto_tracker_update(c->tracker, rgb, c->fw, c->fh, r);
    if (scripted_err) return -1;
    *r = *scripted;
    return 0;
}

int vto_context_process_frame(vto_context* c, const uint8_t* rgb, const vto_result* scripted, int scripted_err, vto_bbox* out) { /* :64-155 */
    vto_result r;
    if (c->state == 0) {
        if (c->pending_confirm) {
            c->pending_confirm = 0;
            if (c->sel.phase == 0) { /* :71-80 */
                c->sel.start_x = c->sel.cursor_x;
                c->sel.start_y = c->sel.cursor_y;
                c->sel.phase = 1;
            } else { /* :81-112 */
                const vto_bbox bb = sel_bbox(&c->sel);
                if (c->tracker) vto_tracker_init(c->tracker, rgb, c->fw, c->fh, bb); /* result ignored, :88 */
                if (ctx_update(c, rgb, scripted, scripted_err, &r) == 0) {
                    if (r.success && r.score > 0.25f) {
                        c->bbox = r.bbox;
                        c->has_bbox = 1;
                        c->score = r.score;
                        c->state = 1;
                        *out = c->bbox;
                        return 1;
                    }
                    sel_new(&c->sel, c->fw, c->fh);
                } else {
                    sel_new(&c->sel, c->fw, c->fh);
                }
            }
        }
        return 0;
    }
    if (c->state == 1) { /* :117-140 */
        c->pending_confirm = 0;
        if (ctx_update(c, rgb, scripted, scripted_err, &r) == 0) {
            if (r.success && r.score > 0.25f) {
                c->bbox = r.bbox;
                c->has_bbox = 1;
                c->score = r.score;
                *out = r.bbox;
                return 1;
            }
            c->state = 2;
            c->lost_frames = 0;
            c->score = 0.f;
            return 0;
        }
        c->state = 2;
        c->lost_frames = 0;
        return 0;
    }
    /* Lost, :142-153 */
    c->pending_confirm = 0;
    if (c->lost_frames > 60) {
        c->state = 0;
        sel_new(&c->sel, c->fw, c->fh);
        c->has_bbox = 0;
    } else {
        c->lost_frames += 1;
    }
    return 0;
}

int vto_context_state(const vto_context* c) { /* :157-166 */
    if (c->state == 0) return c->sel.phase == 0 ? VTO_STATE_SELECT_START : VTO_STATE_SELECT_END;
    return c->state == 1 ? VTO_STATE_TRACKING : VTO_STATE_LOST;
}
const char* vto_context_state_name(const vto_context* c) {
    static const char* names[4] = {"SELECT START", "SELECT END", "TRACKING", "LOST"};
    return names[vto_context_state(c)];
}
float vto_context_score(const vto_context* c) { return c->score; }
int vto_context_bbox(const vto_context* c, vto_bbox* out) {
    if (c->has_bbox) *out = c->bbox;
    return c->has_bbox;
}
void vto_context_selection(const vto_context* c, int32_t o[5]) {
    o[0] = c->sel.cursor_x, o[1] = c->sel.cursor_y, o[2] = c->sel.start_x, o[3] = c->sel.start_y, o[4] = c->sel.phase;
}
uint64_t vto_context_lost_frames(const vto_context* c) { return c->lost_frames; }
