/*
 * livevideo.h -- C ABI of livevideo-b200: the B200-native replacement for the per-frame hot path
 * of wonsang/LiveVideo (HMCPT).  Plain pointers and sizes only; no CUDA or torch types.
 *
 * What each entry replaces in the reference (paths relative to the reference tree):
 *
 *   YV12_to_RGB24          ColorSpaceConversions::YV12_to_RGB24, Convert.h:26 (code: cscc.lib,
 *                          Convert.obj .text+0x350); called at MainDlg.cpp:946,953,960,967,974 and
 *                          BkdextDlg.cpp:106.  Output contract (bottom-up B,G,R DIB, stride 3*w):
 *                          MainDlg.cpp:876-882, 985-995.
 *   FrameDiff_MB           the background / frame difference primitive IsNearlyZero (lib/JM/image.h:22,
 *                          MainDlg.h:9) as applied per pixel in lib/JM/image.c:4791-4804 and
 *                          MainDlg.cpp:804-810, plus the 16x16 macroblock grid and "block is ON" scan of
 *                          lib/JM/image.c:4677-4679, 4981, 5143-5150.
 *   lv_* (context API)     the same two steps fused in one pass over a batch of frames of many
 *                          streams, with the previous-frame retention of lib/JM/image.c:1553-1561
 *                          (prev_frm <- curr_frm) kept as per-stream device state, fed by the planar
 *                          I420 layout JM emits in lib/JM/output.c:427-451.
 *
 * Every entry needs a CUDA device (sm_100a build).  There is no CPU fallback: when no device is
 * usable the int-returning entries return LV_E_NODEVICE / LV_E_CUDA and the two void legacy entries
 * print the CUDA error to stderr and abort().
 */
#ifndef LIVEVIDEO_H
#define LIVEVIDEO_H

#include <stddef.h>
#include <stdint.h>

#if defined(_WIN32)                 /* the reference is a Win32 application: the library is a DLL there */
#ifdef LV_BUILDING_LIBRARY
#define LV_API __declspec(dllexport)
#else
#define LV_API __declspec(dllimport)
#endif
#else
#define LV_API __attribute__((visibility("default")))
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* ---- error codes ------------------------------------------------------------------------------ */
#define LV_OK           0
#define LV_E_ARG       -1   /* NULL / out-of-range argument                                     */
#define LV_E_SIZE      -2   /* unsupported geometry (width % 4 != 0 or height % 2 != 0, ...)    */
#define LV_E_NODEVICE  -3   /* no CUDA device / wrong device index                              */
#define LV_E_CUDA      -4   /* a CUDA runtime call failed; see lv_last_error()                  */
#define LV_E_NOMEM     -5   /* host or device allocation failed                                 */
#define LV_E_STATE     -6   /* call sequence error (e.g. pipeline used before it was created)   */

LV_API const char *lv_strerror(int code);
LV_API const char *lv_last_error(void);      /* thread-local text of the last CUDA failure */
LV_API int lv_version(void);                 /* 10000*major + 100*minor + patch */
LV_API int lv_device_count(void);

/* ---- legacy drop-in entries: host pointers, synchronous, caller owns every buffer ----------------
 * Same parameter list and meaning as Convert.h:26.  src0 = Y (width*height), src1 = U (Cb), src2 = V
 * (Cr) (width*height/4 each, tight stride), dst_ori = 3*width*height bytes, bottom-up rows, B,G,R.
 * Like the original it requires width % 4 == 0 and height % 2 == 0 (the original overruns otherwise);
 * this build checks and aborts instead of overrunning.  Thread-safe; a per-thread context on the
 * current CUDA device is created on first use. */
LV_API void YV12_to_RGB24(unsigned char *src0, unsigned char *src1, unsigned char *src2,
                          unsigned char *dst_ori, int width, int height);

/* mask[i] = (|cur_y[i] - ref_y[i]| > tol) as 0/1 bytes (the reference's bkd_obj, image.c:4798/4803);
 * mb_count[k] = number of set mask pixels in macroblock k = mby*mb_w + mbx, mb_w = ceil(width/16),
 * mb_h = ceil(height/16); mb_map[k] = mb_count[k] > mb_thresh (mb_thresh = 0 is the reference's
 * "any pixel" rule, image.c:5143-5150).  Any output pointer may be NULL. */
LV_API void FrameDiff_MB(const unsigned char *cur_y, const unsigned char *ref_y, int width, int height,
                         int tol, int mb_thresh, unsigned char *mask, unsigned short *mb_count,
                         unsigned char *mb_map);

/* Several pictures per call -- the reference's real call pattern: CMainDlg::DisplayPic converts FIVE pictures back to
 * back (MainDlg.cpp:937-977: pCon.YV12_to_RGB24(srcY[i], srcU[i], srcV[i], RGBbuf[i], width, height), i = 0..4), so the
 * arrays the dialog already holds (unsigned char *srcY[5], ... RGBbuf[5], MainDlg.h) are passed as they are.  The upload
 * of picture i+1, the kernel of picture i and the download of picture i-1 overlap on three streams; the call returns
 * when every dst_ori[i] is complete.  Results are identical to n calls of YV12_to_RGB24 / FrameDiff_MB.  Host
 * pointers (pageable, or pinned in place with lv_host_register for DMA at link speed); returns LV_OK or an LV_E_* code
 * (never aborts).  lv_diff_host_multi: mask / mb_count / mb_map (arrays or single entries) may be NULL; when
 * ref_y[i] == cur_y[i-1] (a sequence differenced against its own previous picture, lib/JM/image.c:1553-1561) every
 * plane is uploaded once. */
LV_API int lv_convert_host_multi(const unsigned char *const *src0, const unsigned char *const *src1,
                                 const unsigned char *const *src2, unsigned char *const *dst_ori, int n, int width,
                                 int height);
LV_API int lv_diff_host_multi(const unsigned char *const *cur_y, const unsigned char *const *ref_y, int n, int width,
                              int height, int tol, int mb_thresh, unsigned char *const *mask,
                              unsigned short *const *mb_count, unsigned char *const *mb_map);

/* The other five methods of ColorSpaceConversions (Convert.h:25-30; code at Convert.obj .text+0x1d0, +0x740, +0x8b0,
 * +0x960, +0xa40).  The application never calls them; they complete the Convert.h surface with the object code's
 * exact semantics (same host-pointer, synchronous, void convention as YV12_to_RGB24; w % 4 == 0, h % 4 == 0):
 *   RGB24_to_YV12: in = bottom-up rows of R,G,B bytes; out = Y, Cb, Cr planes; chroma from the top-left pixel of each
 *                  2x2 block; Y = sum of truncated 0.299/0.587/0.114 products (table-exact)
 *   YVU9_to_YV12 : every input plane is read bottom-up; chroma (w/4 x h/4) replicated 2x2
 *   YUY2_to_YV12 : Y0 U Y1 V rows -> Y plane, then the plane of the V bytes, then the plane of the U bytes; the odd
 *                  row of each row pair supplies the chroma
 *   YV12_to_YVU9 : luma copied, chroma point-sampled 2:1 in both directions
 *   YV12_to_YUY2 : inverse layout of YUY2_to_YV12, one chroma row for both rows of a pair */
LV_API void RGB24_to_YV12(unsigned char *in, unsigned char *out, int w, int h);
LV_API void YVU9_to_YV12(unsigned char *in, unsigned char *out, int w, int h);
LV_API void YUY2_to_YV12(unsigned char *in, unsigned char *out, int w, int h);
LV_API void YV12_to_YVU9(unsigned char *in, unsigned char *out, int w, int h);
LV_API void YV12_to_YUY2(unsigned char *in, unsigned char *out, int w, int h);

/* ---- context API: device pointers, asynchronous on a caller-supplied CUDA stream ------------------ */
typedef struct lv_ctx lv_ctx;

/* One context per (host thread, device).  max_streams = number of independent camera streams whose
 * previous-luma state the context keeps (>= 1).  The state starts as all zero, like the calloc'ed
 * prev_frm of lib/JM/image.c:564-570.  tol in [0,255] (reference default BKD_STR_TOL = 20,
 * lib/JM/global.h:125), mb_thresh in [0,256]. */
LV_API int lv_create(int device, int width, int height, int max_streams, int tol, int mb_thresh,
                     lv_ctx **out);
LV_API void lv_destroy(lv_ctx *ctx);
LV_API int lv_set_params(lv_ctx *ctx, int tol, int mb_thresh);
/* What a frame is differenced against in the fused step.  LV_REF_PREVIOUS (default): frame t-1, the first frame of a
 * batch against the carried previous luma (prev_frm, lib/JM/image.c:1553-1561).  LV_REF_STATIC: every frame of a stream
 * against the stream's stored plane -- the static background BkdBuf of MainDlg.cpp:238-246, set with lv_set_prev_luma --
 * which the step then leaves untouched. */
#define LV_REF_PREVIOUS 0
#define LV_REF_STATIC   1
LV_API int lv_set_ref_mode(lv_ctx *ctx, int mode);

/* Geometry helpers (bytes / element counts per frame). */
LV_API size_t lv_frame_bytes_i420(const lv_ctx *ctx);   /* 1.5*w*h            */
LV_API size_t lv_frame_bytes_bgr(const lv_ctx *ctx);    /* 3*w*h              */
LV_API size_t lv_frame_mask_units(const lv_ctx *ctx);   /* (w/16)*h uint16    */
LV_API size_t lv_frame_mbs(const lv_ctx *ctx);          /* mb_w*mb_h          */

/* Previous-luma state of one stream (checkpoint / stream migration / static background as reference,
 * MainDlg.cpp:238-246).  `y` is a host pointer when is_device == 0, else a device pointer. */
LV_API int lv_set_prev_luma(lv_ctx *ctx, int stream, const uint8_t *y, int is_device, void *cuda_stream);
LV_API int lv_get_prev_luma(lv_ctx *ctx, int stream, uint8_t *y, int is_device, void *cuda_stream);
LV_API int lv_reset_prev_luma(lv_ctx *ctx, void *cuda_stream);   /* back to all zero */

/* The fused step.  d_i420 holds n_streams*n_frames frames, stream-major: frame (s,t) starts at byte
 * (s*n_frames+t)*1.5*w*h and is Y[w*h] U[w*h/4] V[w*h/4] (lib/JM/output.c:427-451).  Stream s maps to
 * state slot first_stream+s.  Frame t=0 is differenced against the stream's carried previous luma,
 * frame t>0 against frame t-1 of the batch; afterwards the state holds frame n_frames-1's luma.
 * Outputs (device pointers; each may be NULL to skip it):
 *   d_bgr        3*w*h bytes per frame, bottom-up B,G,R (identical to YV12_to_RGB24)
 *   d_mask_bits  (w/16)*h uint16 per frame; bit i of unit u of row y = mask of pixel (16u+i, y)
 *   d_mask_bytes w*h bytes per frame, 0/1
 *   d_mb_count   mb_w*mb_h uint16 per frame;  d_mb_map  mb_w*mb_h bytes per frame
 *   d_summary    2 uint32 per frame: {changed pixels, changed macroblocks}
 * Requires width % 16 == 0 and height % 2 == 0.
 * Alignment (checked, LV_E_ARG otherwise -- the kernels move 16-byte words, and a misaligned vector access would be a
 * sticky CUDA error): d_i420, d_bgr, d_mask_bytes 16 bytes; d_mask_bits, d_mb_count 2 bytes; d_summary 4 bytes.  The same
 * holds for lv_convert_batch (d_i420, d_bgr) and lv_diff_batch (d_cur_y, d_ref_y, d_mask_bytes 16 bytes).
 * lv_write_bks needs 4-byte aligned frames (16-byte aligned ones take the vector kernel); lv_narrow_pictures (beyond the
 * 2 bytes of its uint16 input) and lv_convert_format accept any byte alignment (they fall back to byte accesses).  Anything from cudaMalloc / a torch allocation is 256-byte aligned; only sub-buffers at odd offsets are
 * affected.
 * Streams keep their state independently: stepping streams [first_stream, first_stream + n_streams) neither reads nor
 * copies the state of any other stream. */
LV_API int lv_convert_diff_batch(lv_ctx *ctx, const uint8_t *d_i420, int first_stream, int n_streams,
                                 int n_frames, uint8_t *d_bgr, uint16_t *d_mask_bits,
                                 uint8_t *d_mask_bytes, uint16_t *d_mb_count, uint8_t *d_mb_map,
                                 uint32_t *d_summary, void *cuda_stream);

/* Same, with flags.  LV_STEP_SUMMARY_PREZEROED: d_summary already holds zeros (the blocks of the fused kernel add into
 * it; the plain entry issues a memset in front of the launch). */
#define LV_STEP_SUMMARY_PREZEROED 1
LV_API int lv_convert_diff_batch_ex(lv_ctx *ctx, const uint8_t *d_i420, int first_stream, int n_streams,
                                    int n_frames, uint8_t *d_bgr, uint16_t *d_mask_bits, uint8_t *d_mask_bytes,
                                    uint16_t *d_mb_count, uint8_t *d_mb_map, uint32_t *d_summary, void *cuda_stream,
                                    int flags);

/* The fused step with the step AFTER the path folded in: the object box of motion_segment's I-frame branch
 * (lib/JM/image.c:4783-4848).  For each frame, n_objects windows {x0, x1, y0, y1} (inclusive luma coordinates; the
 * reference's min_x, max_x, min_y, max_y of image.c:4786-4789) and, out, n_objects boxes {min_x, max_x, min_y, max_y} =
 * the extreme coordinates of the changed pixels (bkd_obj == 1, image.c:4798-4803) inside the window (image.c:4818-4838);
 * a window without a changed pixel yields the reference's initial values {x1+1, x0-1, y1+1, y0-1}.  The boxes are
 * accumulated by the fused kernel itself (warp min/max reductions, at most four atomics per block and object): no mask
 * is written to or read back from memory.  With lv_set_ref_mode(LV_REF_STATIC) and the background in the state
 * (lv_set_prev_luma) this is the reference's background subtraction against BkdBuf; d_bgr may be NULL (the reference's
 * step displays nothing).  d_windows / d_boxes: 4 ints per (frame, object), frame-major, 16-byte aligned.
 *   lv_object_windows  windows from objects {PosX, PosY, Width, Height}: ClipX(PosX -+ Width/2 -+ hor_margin), ClipY(...)
 *                      (image.c:4786-4789, image.h:23-24; the reference's margins are 8, global.h:119-120)
 *   lv_object_update   objects from boxes under the reference's guard (image.c:4840-4848): only when min_x < max_x and
 *                      min_y < max_y: PosX = min_x + (max_x-min_x)/2, PosY likewise, Width = max_x-min_x, Height likewise */
#define LV_MAX_OBJECTS 16
LV_API int lv_convert_diff_box_batch(lv_ctx *ctx, const uint8_t *d_i420, int first_stream, int n_streams, int n_frames,
                                     uint8_t *d_bgr, uint16_t *d_mb_count, uint8_t *d_mb_map, uint32_t *d_summary,
                                     const int *d_windows, int n_objects, int *d_boxes, void *cuda_stream);
/* Same, with flags.  LV_BOX_WINDOWED: the boxes come from a pass of their own over the WINDOWS only, behind the plain fused step
 * (which then runs at its full speed), re-reading current and reference luma inside each window: 117.4 us instead of 119.4 us
 * per 64 pictures of 1080p with one window of the reference's size per picture, 119.5 instead of 121.0 with four -- the pass is a
 * small dependent kernel, about 5 us whatever it does.  For windows that cover a small part of the picture, as the reference's
 * do (an object's last box plus a margin, lib/JM/image.c:4786-4789); a window as large as the picture costs a second, slow pass
 * over the luma (340 us), which is why the fused form is the default.  Results are identical.  LV_STEP_SUMMARY_PREZEROED as for
 * lv_convert_diff_batch_ex. */
#define LV_BOX_WINDOWED 2
LV_API int lv_convert_diff_box_batch_ex(lv_ctx *ctx, const uint8_t *d_i420, int first_stream, int n_streams, int n_frames,
                                        uint8_t *d_bgr, uint16_t *d_mb_count, uint8_t *d_mb_map, uint32_t *d_summary,
                                        const int *d_windows, int n_objects, int *d_boxes, void *cuda_stream, int flags);
LV_API int lv_object_windows(lv_ctx *ctx, const int *d_objects, int n, int hor_margin, int ver_margin, int *d_windows,
                             void *cuda_stream);
LV_API int lv_object_update(lv_ctx *ctx, const int *d_boxes, int n, int *d_objects, void *cuda_stream);

/* Conversion only (K2) and difference only (K3) over n frames; same layouts as above.
 * lv_diff_batch: frame i of d_cur_y (w*h bytes each) against frame i of d_ref_y. */
LV_API int lv_convert_batch(lv_ctx *ctx, const uint8_t *d_i420, int n_frames, uint8_t *d_bgr,
                            void *cuda_stream);
LV_API int lv_diff_batch(lv_ctx *ctx, const uint8_t *d_cur_y, const uint8_t *d_ref_y, int n_frames,
                         uint16_t *d_mask_bits, uint8_t *d_mask_bytes, uint16_t *d_mb_count,
                         uint8_t *d_mb_map, uint32_t *d_summary, void *cuda_stream);

/* ---- the steps beside / after the fused path that the application shows ----------
 * lv_write_bks: the literal CMainDlg::WriteBks (MainDlg.cpp:771-835): for every sample of n_frames I420 frames,
 *   out = IsNearlyZero(cur - bkd, tol) ? fill : cur, fill = 16 (Y) / 128 (U, V), against ONE background frame
 *   d_bkd_i420 (MainDlg.cpp:238-246).  d_obj_bits (optional, width % 16 == 0) receives bkd_obj
 *   (lib/JM/image.c:4798,4803) as a luma bit mask in the d_mask_bits layout.  tol in [0,255] (the reference uses 25
 *   here, MainDlg.cpp:780, and BKD_STR_TOL = 20 in image.c:4698).
 * lv_object_box: min/max x,y of the set mask pixels inside the window [x0..x1] x [y0..y1] (inclusive) of ONE
 *   frame (lib/JM/image.c:4818-4838); d_box4 = {min_x, max_x, min_y, max_y}; an empty window gives
 *   {x1+1, x0-1, y1+1, y0-1}, the reference's initial values.  mask_is_bits selects the d_mask_bits layout. */
LV_API int lv_write_bks(lv_ctx *ctx, const uint8_t *d_cur_i420, const uint8_t *d_bkd_i420, int n_frames, int tol,
                        uint8_t *d_out_i420, uint16_t *d_obj_bits, void *cuda_stream);
LV_API int lv_object_box(lv_ctx *ctx, const void *d_mask, int mask_is_bits, int x0, int x1, int y0, int y1,
                         int *d_box4, void *cuda_stream);
/* The same box for each of n_frames masks laid out back to back (the d_mask_bytes / d_mask_bits outputs of a fused
 * launch) in ONE pass: d_boxes = 4 ints per frame.  d_boxes (and a byte mask when width % 16 == 0) must be 16-byte
 * aligned. */
LV_API int lv_object_box_batch(lv_ctx *ctx, const void *d_mask, int mask_is_bits, int n_frames, int x0, int x1, int y0,
                               int y1, int *d_boxes, void *cuda_stream);

/* The GPU side of the producer hook.  write_out_picture (lib/JM/output.c:362-453) narrows the
 * decoder's unsigned-short planes (imgpel, lib/JM/global.h:61) to bytes and cuts the crop rectangle away with one
 * memcpy per SAMPLE (img2buf, output.c:161-176: the low byte of the little-endian short, no clipping).  Here the host
 * only block-copies the raw planes (each is one contiguous allocation, memalloc.c:26-39) and the device does the rest:
 * d_pic16 = n_frames pictures, each Y16[size_x*size_y] U16[(size_x/2)*(size_y/2)] V16[...] (U = imgUV[0], V = imgUV[1]);
 * d_i420 = n_frames tight I420 frames of the context's geometry, whose rectangle starts at (crop_left, crop_top) of the
 * picture (luma pixels, even: twice the SPS offsets, output.c:375-380; the chroma planes use half, output.c:421-425). */
LV_API int lv_narrow_pictures(lv_ctx *ctx, const uint16_t *d_pic16, int n_frames, int size_x, int size_y,
                              int crop_left, int crop_top, uint8_t *d_i420, void *cuda_stream);

/* The same pictures straight into the fused step: lv_narrow_pictures + lv_convert_diff_batch in ONE kernel (the low byte of
 * every sample is taken and the crop cut away while the pictures are read, so the I420 frames never exist in memory:
 * 3 B/px in + 3 B/px out instead of 4.5 + 4.5; frame t > 0 is differenced against the previous PICTURE's luma).  d_pic16 =
 * n_streams * n_frames pictures, stream-major, in the lv_narrow_pictures layout; outputs, state and results exactly as
 * lv_convert_diff_batch (d_bgr is mandatory here).  The one-kernel path needs d_pic16 16-byte aligned, size_x % 16 == 0,
 * crop_left % 16 == 0 and plane sizes that are multiples of 8 samples (true for every macroblock-aligned picture); any
 * other legal layout is served by the two kernels through a scratch the context owns. */
LV_API int lv_convert_diff_pictures(lv_ctx *ctx, const uint16_t *d_pic16, int size_x, int size_y, int crop_left, int crop_top,
                                    int first_stream, int n_streams, int n_frames, uint8_t *d_bgr, uint16_t *d_mask_bits,
                                    uint8_t *d_mask_bytes, uint16_t *d_mb_count, uint8_t *d_mb_map, uint32_t *d_summary,
                                    void *cuda_stream);

/* Device-pointer form of the five converters above: kind 1 RGB24_to_YV12, 2 YVU9_to_YV12, 3 YUY2_to_YV12,
 * 4 YV12_to_YVU9, 5 YV12_to_YUY2; n_frames frames back to back (bytes per frame: lv_format_bytes). */
LV_API int lv_convert_format(lv_ctx *ctx, int kind, const uint8_t *d_in, uint8_t *d_out, int n_frames, void *cuda_stream);
LV_API size_t lv_format_bytes(int kind, int width, int height, int output);

/* ---- host-buffer step: pinned, double-buffered H2D -> fused kernel -> D2H ------------------------
 * The call a host application makes when frames live in host memory (the JM producer hook of
 * lib/JM/output.c:427-451 writes into the pinned input).  lv_host_alloc / lv_host_free hand out
 * pinned memory; lv_step_host splits the batch into chunks of `chunk_frames` frames per stream group
 * and overlaps the upload of chunk k+1 (copy stream) with the kernel of chunk k and the download of
 * chunk k-1.  The first chunks of a step are small (1, 2, 4 ... frames, at least 1 MiB of input) so that the download
 * stream -- the bottleneck: 3 of the 4.5 bytes per pixel that cross the link -- starts after one small upload; the
 * macroblock counts, the map and the summaries are staged for the whole step on the device and downloaded once, beside
 * the last pictures.  Host pointers may also be ordinary pageable memory (slower: staged by the driver).
 * Outputs may be NULL.  Synchronous: returns when every output is in host memory.
 * Environment (read at every call): LV_HOST_CHUNK_MB = input MiB per chunk (default 16), LV_HOST_SLOT_MB = MiB of input
 * one of the three device slots may hold (default 256), LV_HOST_RAMP=0 switches the small first chunks off. */
/* The upload half on its own, for consumers that keep the BGR on the device: lv_upload_async copies n_frames_total
 * I420 frames from (pinned) host memory into the context's device input slot `slot` (0 or 1, sized on demand) on the
 * context's copy stream and returns the device pointer; lv_upload_wait makes `cuda_stream` wait for that slot's copy.
 * Alternating slots double-buffers: upload batch k+1 into slot (k+1)&1 while the kernel of batch k reads slot k&1.
 * A slot must not be re-uploaded before the kernel that read it has been enqueued-and-finished (lv_upload_async
 * waits for the event the caller records with lv_upload_release). */
LV_API int lv_upload_async(lv_ctx *ctx, const uint8_t *h_i420, int n_frames_total, int slot, const uint8_t **d_i420);
LV_API int lv_upload_wait(lv_ctx *ctx, int slot, void *cuda_stream);
LV_API int lv_upload_release(lv_ctx *ctx, int slot, void *cuda_stream);   /* call after enqueuing the slot's consumer */
LV_API int lv_host_alloc(void **p, size_t bytes);
LV_API int lv_host_free(void *p);
/* Pin buffers the application already owns (the reference mallocs its planes and RGB buffer once and reuses them for
 * every picture, MainDlg.cpp:39-51): every entry that takes host pointers -- the legacy ones included -- then copies
 * by DMA instead of through the driver's pageable staging.  Keep the range allocated until lv_host_unregister. */
LV_API int lv_host_register(void *p, size_t bytes);
LV_API int lv_host_unregister(void *p);
/* 1 when [p, p + bytes) lies in a range pinned AND mapped through this library, i.e. the legacy entries take the zero-copy path
 * for it; 0 otherwise (pageable, or pinned elsewhere).
 * Environment LV_AUTO_PIN=1 (opt-in, read once): the legacy entries YV12_to_RGB24 / FrameDiff_MB register, in place, every
 * buffer they see for the second time with the same address and size -- the effect of one lv_host_register per buffer for an
 * application that is relinked without a source change (1080p: 589 -> 188 us per YV12_to_RGB24 call from the third call on).
 * Only for applications that keep their buffers for their lifetime, as the reference does (MainDlg.cpp:39-51): a range stays
 * registered until the process ends. */
LV_API int lv_host_is_mapped(const void *p, size_t bytes);
/* How lv_step_host / lv_step_host_pictures would cut a step of n_streams x n_frames frames into chunks under the current
 * environment, without touching a device (pure host logic; the CPU tests walk it): returns the number of chunks and stores
 * up to max_chunks of them as {first stream, streams, first frame, frames} in chunks4.  in_frame_bytes = host bytes per
 * frame (0: I420, width * height * 3 / 2; decoder pictures: 2 bytes per sample). */
LV_API int lv_host_step_plan(int width, int height, size_t in_frame_bytes, int n_streams, int n_frames, int *chunks4,
                             int max_chunks);
LV_API int lv_step_host(lv_ctx *ctx, const uint8_t *h_i420, int first_stream, int n_streams, int n_frames,
                        uint8_t *h_bgr, uint16_t *h_mask_bits, uint16_t *h_mb_count, uint8_t *h_mb_map,
                        uint32_t *h_summary);
/* The same step fed with the decoder's own pictures instead of I420 frames: h_pic16 holds n_streams*n_frames pictures
 * (stream-major) in the lv_narrow_pictures layout; the narrowing + cropping of lib/JM/output.c:87-178, 362-453 runs on
 * the device in front of the fused kernel of every chunk, so the producer hook is three block copies per picture. */
LV_API int lv_step_host_pictures(lv_ctx *ctx, const uint16_t *h_pic16, int size_x, int size_y, int crop_left, int crop_top,
                                 int first_stream, int n_streams, int n_frames, uint8_t *h_bgr, uint16_t *h_mask_bits,
                                 uint16_t *h_mb_count, uint8_t *h_mb_map, uint32_t *h_summary);

/* ---- synthetic input (benchmark / tests): stands in for the JM decoder's output --------------------
 * kind 0 = `uniform`, 1 = `camera`; writes n_streams*n_frames I420 frames, stream-major,
 * for streams first_stream.. and frames first_frame.. into device memory.  Byte-identical to the same
 * generator stated in C in oracle/lv_oracle.c. */
LV_API int lv_gen_synthetic(int kind, uint32_t seed, uint32_t first_stream, uint32_t first_frame, int n_streams,
                            int n_frames, int width, int height, uint8_t *d_i420, void *cuda_stream);

/* ---- multi-GPU from one host process: streams sharded by stream ID, no pixel ever crosses GPUs ---------------
 * Device g (devices[g], or g when devices == NULL) owns the contiguous block of streams [g*S/G, (g+1)*S/G)
 * (lv_shard_range).  lv_shard_step only enqueues one fused launch per device on that device's own stream, so the
 * devices run concurrently; d_i420[g] / d_bgr[g] / d_mb_count[g] / d_mb_map[g] are pointers in device g's memory
 * holding that block (same layouts as lv_convert_diff_batch; output arrays or entries may be NULL).
 * lv_gather_summaries is the only exchange of the path: an NCCL all-gather (bound at run time from libnccl.so.2) of
 * {changed_px, changed_mbs} per (stream, frame) of the LAST step -- every device ends with the whole table
 * (lv_shard_gathered, padded per rank to the largest block) and, when h_summary != NULL, the unpadded
 * [total_streams][n_frames][2] table is copied to the host (synchronises device 0's stream only).
 * With one device no collective is issued.  Returns LV_E_STATE when NCCL is needed but cannot be loaded. */
typedef struct lv_shard lv_shard;
LV_API int lv_shard_create(int n_devices, const int *devices, int width, int height, int total_streams, int tol,
                           int mb_thresh, lv_shard **out);
LV_API void lv_shard_destroy(lv_shard *s);
LV_API int lv_shard_device_count(const lv_shard *s);
LV_API int lv_shard_range(const lv_shard *s, int rank, int *device, int *first_stream, int *n_streams);
LV_API lv_ctx *lv_shard_ctx(lv_shard *s, int rank);        /* rank's own context (state get/set, other entries) */
LV_API void *lv_shard_stream(lv_shard *s, int rank);       /* rank's cudaStream_t */
LV_API int lv_shard_step(lv_shard *s, const uint8_t *const *d_i420, int n_frames, uint8_t *const *d_bgr,
                         uint16_t *const *d_mb_count, uint8_t *const *d_mb_map);
LV_API int lv_gather_summaries(lv_shard *s, uint32_t *h_summary);
LV_API const uint32_t *lv_shard_gathered(lv_shard *s, int rank);
LV_API int lv_shard_sync(lv_shard *s);

/* ---- kernel variant (process-wide) -------------------------------------------------------------------
 * 0 = tiled kernel (default), 1 = TMA pipeline (cp.async.bulk.tensor loads/stores; measured slower, kept as the one
 * selectable alternative), 2 / 3 = tiled kernel with its chroma loads forced late / early (0 picks by row length: early up
 * to 2048 pixels per row).  All are bit-identical; the choice only affects speed (DESIGN.md has the measurements). */
LV_API int lv_set_kernel_variant(int variant);
LV_API int lv_get_kernel_variant(void);

/* ---- zero-copy mode of the host-pointer conversion entries (process-wide) ------------------------------
 * Buffers handed out by lv_host_alloc or pinned in place by lv_host_register are also MAPPED into the device's address
 * space.  When every buffer of a YV12_to_RGB24 / lv_convert_host_multi / FrameDiff_MB call is such a buffer (planes and
 * picture 16-byte aligned, chroma planes 8-byte), the kernel reads the planes from host memory and writes the result to
 * host memory itself: one launch + one synchronise per call, upload and download of the same picture overlapped on the
 * link.  0 = off (always stage through device buffers), 1 = on with the tile kernel's store pattern, 2 = on with stores
 * re-dealt to 512 contiguous bytes per warp instruction (default; environment variable LV_ZEROCOPY sets the initial
 * value).  Results are identical in every mode.  Replaces nothing in the reference's interface: the calls stay
 * Convert.h:26 / MainDlg.cpp:946-974 as they are. */
LV_API int lv_set_zero_copy(int mode);
LV_API int lv_get_zero_copy(void);

/* ---- introspection for the benchmark ------------------------------------------------------------- */
LV_API uint64_t lv_kernel_launches(const lv_ctx *ctx);   /* kernels of this library launched so far */
LV_API int lv_device_sm_count(const lv_ctx *ctx);

/* ---- peer-memory exchange of the change summaries (one process per GPU; csrc/lv_exchange.cu) ------------------------
 * The only exchange of the sharded path: every rank wants every rank's {changed_px, changed_mbs} per
 * (stream, frame).  Instead of a per-step library all-gather, each rank pushes its block into every peer's table with
 * NVLink stores -- every 8-byte store carries one word and its own sequence number, so there is no flag and no fence --
 * and a collect spins on local memory only.  Ranks never rendezvous; a rank waits only when it is about half a ring
 * (`depth` / 2 steps) ahead of a peer.
 *
 *   lv_exchange_create     allocates this rank's buffer (ring of `depth` tables of world x slot_words words)
 *   lv_exchange_handle     64-byte CUDA IPC handle of it; hand the handles of all ranks to lv_exchange_connect
 *                          (any transport: torch.distributed, MPI, a file) or pass already-mapped pointers to
 *                          lv_exchange_connect_ptrs (one process driving several ranks, VMM / symmetric allocations)
 *   lv_exchange_step       the whole sharded step: the fused kernel accumulates into a ring row owned by the exchange;
 *                          behind an event, on the exchange's own side stream, one small kernel pushes that row to
 *                          every peer (and copies it to d_summary, optional) and another collects an older step.  The
 *                          caller's stream receives the two launches of the single-GPU step plus one event record.
 *   lv_exchange_result     table of `step` (world rows of slot_words u32, device memory owned by the exchange, valid
 *                          until `depth` further steps have been issued); blocks the host until every rank's block of
 *                          that step has arrived (every rank must have issued that step); afterwards d_summary of
 *                          every step up to `step` is filled too
 *   lv_exchange_flush      nothing to do in this design (every push is on its way when lv_exchange_step returns);
 *                          kept so that callers need not know
 *   lv_exchange_publish / lv_exchange_collect   the two halves on their own (not to be mixed with lv_exchange_step):
 *                          publish(step, d_block, n_words) in order, once per step; collect(step, d_table) in order,
 *                          within `depth` steps; d_table = world x slot_words u32 or NULL to retire the step unread
 *   lv_exchange_error      0, or which wait gave up (1 publish / 2 collect / 3 gate): a kernel of the exchange never
 *                          spins for more than 4 s in all (one deadline per kernel), and once the error is set every
 *                          later kernel returns at once; read from a host-mapped word, no copy, no synchronisation
 *   lv_exchange_gate       device-side rendezvous of all ranks in the caller's stream (collective: same number of calls
 *                          on every rank): what is enqueued behind it starts within microseconds of the same instant on
 *                          every GPU, however far apart the host threads are.  hold != 0: this rank announces its arrival
 *                          only after lv_exchange_gate_open -- call that once the work behind the gate is enqueued, and
 *                          when the gate opens every GPU has its work in the queue (no GPU starts late because its host
 *                          was the last to launch)
 *   lv_exchange_wait_summary  makes a stream wait until the caller's d_summary of `step` has been written: with
 *                          lv_exchange_step that buffer is filled by the push kernel on the exchange's own stream (NOT
 *                          in-stream behind the fused kernel as in lv_convert_diff_batch); a consumer on the compute
 *                          stream must call this first, or use the table of lv_exchange_result
 *   lv_exchange_destroy    only after every rank has collected the last step it will publish (a barrier of the
 *                          launcher): peers store into this rank's buffer until then */
typedef struct lv_exchange lv_exchange;
LV_API size_t lv_exchange_bytes(int world, int depth, int slot_words);
LV_API int lv_exchange_create(int device, int rank, int world, int depth, int slot_words, lv_exchange **out);
LV_API void lv_exchange_destroy(lv_exchange *x);
LV_API void *lv_exchange_buffer(lv_exchange *x);
LV_API int lv_exchange_handle(lv_exchange *x, unsigned char *handle64);
LV_API int lv_exchange_connect(lv_exchange *x, const unsigned char *handles);
LV_API int lv_exchange_connect_ptrs(lv_exchange *x, void *const *peer_buffers);
LV_API int lv_exchange_step(lv_exchange *x, lv_ctx *ctx, const uint8_t *d_i420, int first_stream, int n_streams,
                            int n_frames, uint8_t *d_bgr, uint16_t *d_mask_bits, uint8_t *d_mask_bytes,
                            uint16_t *d_mb_count, uint8_t *d_mb_map, uint32_t *d_summary, void *cuda_stream,
                            uint64_t *step);
LV_API int lv_exchange_flush(lv_exchange *x);
LV_API int lv_exchange_result(lv_exchange *x, uint64_t step, const uint32_t **d_table);
LV_API int lv_exchange_publish(lv_exchange *x, uint64_t step, const uint32_t *d_block, int n_words, void *cuda_stream);
LV_API int lv_exchange_collect(lv_exchange *x, uint64_t step, uint32_t *d_table, void *cuda_stream);
LV_API int lv_exchange_error(lv_exchange *x);
LV_API int lv_exchange_gate(lv_exchange *x, void *cuda_stream, int hold);
LV_API int lv_exchange_gate_open(lv_exchange *x);
/* diagnostics: host-mapped GPU time stamps (ns) of the newest step's push / collect kernels; see csrc/lv_exchange.cu */
LV_API const unsigned long long *lv_exchange_trace(lv_exchange *x);
LV_API int lv_exchange_wait_summary(lv_exchange *x, uint64_t step, void *cuda_stream);
/* device-side form of lv_exchange_result: `cuda_stream` waits until the table of `step` is complete (no host round
 * trip: enqueue the kernel that consumes the table behind it); *d_table (optional) as in lv_exchange_result */
LV_API int lv_exchange_wait_table(lv_exchange *x, uint64_t step, void *cuda_stream, const uint32_t **d_table);
LV_API const char *lv_exchange_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* LIVEVIDEO_H */
