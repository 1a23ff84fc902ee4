/*
 * lv_oracle.h -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the reference's per-frame hot path (wonsang/LiveVideo):
 *   - YUV 4:2:0 -> 24-bit RGB conversion: class ColorSpaceConversions, /root/reference/Convert.h:19-31,
 *     code only as object code in /root/reference/cscc.lib (Convert.obj .text+0x000 ctor tables,
 *     .text+0x350 YV12_to_RGB24).
 *   - background/frame difference: IsNearlyZero, /root/reference/lib/JM/image.h:22 (= MainDlg.h:9),
 *     loops lib/JM/image.c:4791-4804 and MainDlg.cpp:804-810.
 *   - 16x16 macroblock grid and "block is ON" scan: lib/JM/image.c:4677-4679, 4981, 5143-5150.
 *   - previous-frame retention: lib/JM/image.c:1553-1561 (alloc 556-570, calloc => zero initial state).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * link or call this library.  The product (livevideo_b200/csrc) never does.
 *
 * Parity status: the conversion is PINNED -- tests check it against the reference's original
 * object code (oracle/_ref/oracle32, built by oracle/ref32/build_ref.py) on all 2^24 (Y,U,V)
 * triples and against the golden digests of SURVEY.md section 8(c).  The difference primitive
 * (lvo_is_nearly_zero) and img2buf are PINNED against the reference's own C as compiled into
 * oracle/_ref/jm_ldecod (--nearlyzero: all 511 x 256 pairs; --img2buf), tests/golden/*.json.  The
 * loops around the primitive and the macroblock scan have no upstream test or callable binary:
 * "parity unpinned" by the reference, pinned by the cited source lines only.
 */
#ifndef LV_ORACLE_H
#define LV_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- conversion (Convert.obj) ------------------------------------------------------------- */
void lvo_init_tables(void);                                   /* ctor, .text+0x110..0x1c1 */
void lvo_yv12_to_rgb24(const uint8_t *src_y, const uint8_t *src_u, const uint8_t *src_v,
                       uint8_t *dst, int width, int height);  /* .text+0x350..0x73f */
void lvo_yuv_to_bgr_px(int y, int u, int v, uint8_t bgr[3]);  /* one pixel through the tables */
/* textbook >>16 form -- NOT the reference; only used by tests to show the two differ */
void lvo_textbook_px(int y, int u, int v, uint8_t bgr[3]);

/* ---- difference + macroblock reduction ---------------------------------------------------- */
int  lvo_is_nearly_zero(int a, int tol);                      /* image.h:22 */
/* mask[i] = !IsNearlyZero(cur[i]-ref[i], tol); mb_count[k] = sum of mask over MB k
 * (k = mby*mb_w + mbx, mb_w = ceil(w/16), mb_h = ceil(h/16)); mb_map[k] = mb_count[k] > mb_thresh.
 * Any output pointer may be NULL.  Returns changed pixel count. */
uint32_t lvo_frame_diff_mb(const uint8_t *cur_y, const uint8_t *ref_y, int width, int height,
                           int tol, int mb_thresh, uint8_t *mask, uint16_t *mb_count, uint8_t *mb_map);
/* pack a 0/1 byte mask into LSB-first bits, one uint16 per 16-pixel unit (width % 16 == 0) */
void lvo_pack_mask16(const uint8_t *mask, int width, int height, uint16_t *bits);

/* ---- the fused batch step (what the GPU kernel computes) ---------------------------------- */
/* i420: n_streams*n_frames frames, stream-major, each Y[w*h] U[w*h/4] V[w*h/4].
 * prev_luma: n_streams planes; frame 0 of a stream is differenced against it, frame t against
 * frame t-1 of the batch; on return it holds the last frame's luma (image.c:1553-1561).
 * bgr: 3*w*h per frame.  mask (bytes, optional), mb_count, mb_map, summary[2] per frame
 * ({changed_px, changed_mbs}) -- any may be NULL.  n_threads<=1 => single thread. */
void lvo_convert_diff_batch(const uint8_t *i420, uint8_t *prev_luma, int n_streams, int n_frames,
                            int width, int height, int tol, int mb_thresh,
                            uint8_t *bgr, uint8_t *mask, uint16_t *mb_count, uint8_t *mb_map,
                            uint32_t *summary, int n_threads);

/* ---- background subtraction, the literal WriteBks (MainDlg.cpp:771-835) ------------------- */
/* out planes: Y -> 16 where |d|<=tol else cur; U,V -> 128 where |d|<=tol else cur. */
void lvo_write_bks(const uint8_t *cur_i420, const uint8_t *bkd_i420, int width, int height,
                   int tol, uint8_t *out_i420);

/* ---- the other five converters of cscc.lib (Convert.h:25-30; never called by the application) -----------------
 * Restated from the object code: RGB24_to_YV12 .text+0x1d0 (tables built by the ctor at .text+0x025..0x10a),
 * YVU9_to_YV12 +0x740, YUY2_to_YV12 +0x8b0, YV12_to_YVU9 +0x960, YV12_to_YUY2 +0xa40.  "YV12" in these four
 * packed/planar shuffles is the real YV12 (Y, then the plane fed by YUY2's V bytes, then U's); RGB24_to_YV12 writes
 * Y, Cb, Cr.  Pinned against oracle32 modes 1..5 by tests/test_oracle.py. */
void lvo_rgb24_to_yv12(const uint8_t *in, uint8_t *out, int w, int h);   /* in: bottom-up rows, bytes R,G,B */
void lvo_yvu9_to_yv12(const uint8_t *in, uint8_t *out, int w, int h);    /* in is bottom-up; 2x2 replication */
void lvo_yuy2_to_yv12(const uint8_t *in, uint8_t *out, int w, int h);    /* chroma of the odd rows wins */
void lvo_yv12_to_yvu9(const uint8_t *in, uint8_t *out, int w, int h);    /* chroma point-sampled 2:1 both ways */
void lvo_yv12_to_yuy2(const uint8_t *in, uint8_t *out, int w, int h);

/* ---- the decoder's picture -> planar 8-bit I420 (lib/JM/output.c) --------------------------------------------
 * lvo_img2buf: img2buf (output.c:87-178), the little-endian branch with symbol_size_in_bytes == 1 and
 * imgpel == unsigned short (global.h:61): every sample keeps ONE byte of its little-endian short -- the low byte, no
 * clipping -- and the crop rectangle is cut away.  Pinned against the reference's own compiled img2buf
 * (oracle/_ref/jm_ldecod --img2buf, tests/test_oracle.py, tests/golden).
 * lvo_picture16_to_i420: write_out_picture's plane order and crops for 4:2:0 (output.c:375-385, 427-451): Y with the
 * luma crop, then U (imgUV[0]) and V (imgUV[1]) with half of it.  pic16 = Y16[size_x*size_y] U16 V16 (size_x/2 x
 * size_y/2 each), each plane contiguous like get_mem2Dpel's (memalloc.c:26-39).  Crops in luma pixels (even). */
void lvo_img2buf(const uint16_t *plane, int size_x, int size_y, int crop_left, int crop_right, int crop_top,
                 int crop_bottom, uint8_t *buf);
void lvo_picture16_to_i420(const uint16_t *pic16, int size_x, int size_y, int crop_left, int crop_right, int crop_top,
                           int crop_bottom, uint8_t *i420);

/* ---- object box from the mask (image.c:4818-4848) ----------------------------------------- */
/* box[4] = {min_x, max_x, min_y, max_y} over mask==1 inside the window [x0..x1]x[y0..y1] (inclusive);
 * empty => {x1+1, x0-1, y1+1, y0-1} exactly as the reference initialises them. */
void lvo_object_box(const uint8_t *mask, int width, int x0, int x1, int y0, int y1, int box[4]);
/* lib/JM/image.c:4786-4789 (window of an object, ClipX/ClipY of image.h:23-24) and 4840-4848 (object from its box) */
void lvo_object_window(const int obj[4], int width, int height, int hor_margin, int ver_margin, int win[4]);
void lvo_object_update(const int box[4], int obj[4]);

/* ---- synthetic generators; counter based, identical in C / numpy / CUDA --- */
uint32_t lvo_lowbias32(uint32_t x);
void lvo_gen_uniform(uint32_t seed, uint32_t stream, uint32_t frame, int width, int height, uint8_t *i420);
void lvo_gen_camera(uint32_t seed, uint32_t stream, uint32_t frame, int width, int height, uint8_t *i420);

uint32_t lvo_crc32(const uint8_t *p, uint64_t n);

#ifdef __cplusplus
}
#endif
#endif
