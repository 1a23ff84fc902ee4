/*
 * jm_inproc.c -- the producer hook of BASELINE.json configs[4], IN PROCESS: the reference's JM 10.2 decoder (compiled
 * unmodified from /root/reference/lib/JM by oracle/jm/build_jm.py) and liblivevideo.so live in one process, and every
 * decoded picture goes from the decoder's own planes to the GPU without passing through a file, a pipe or img2buf.
 *
 *   jm_inproc stream.264 width height batch out_prefix
 *
 * The decoder hands every output picture to write_out_picture(StorablePicture *p, int p_out) (lib/JM/output.c:362-453;
 * callers write_picture / write_unpaired_field / flush_direct_output in the same file).  The reference's version narrows
 * the unsigned-short planes (imgpel, lib/JM/global.h:61) to bytes with one memcpy per SAMPLE (img2buf, output.c:87-178)
 * and write()s them.  Here that symbol is replaced at link time -- the reference's object file is linked with its own
 * write_out_picture weakened (objcopy --weaken-symbol; no source is edited) -- by the function below, which
 *   1. block-copies the three planes, each one contiguous allocation (get_mem2Dpel, lib/JM/memalloc.c:26-39:
 *      p->imgY[0], p->imgUV[0][0], p->imgUV[1][0]), into a pinned batch of uint16 pictures, and
 *   2. every `batch` pictures calls lv_step_host_pictures: narrowing + cropping + conversion + difference on the GPU.
 * Outputs (BGR, macroblock counts, block map, summaries of all frames) are written to out_prefix.{bgr,cnt,map,sum};
 * one JSON line with the timings goes to stdout.  A maintainer's binding in the application is the same function
 * (INTEGRATION.md).
 */
#include <pthread.h>
#include <stdio_ext.h>
#include <time.h>

#include "jm_driver.h"          /* oracle/jm: the decoder driver (reference headers) */
#include "livevideo.h"

extern FILE *bits;              /* lib/JM/annexb.c:26 */

static struct {
    lv_ctx *ctx;
    int width, height;          /* cropped geometry the context was made for */
    int batch, filled;
    int size_x, size_y, crop_left, crop_top;
    size_t samples;             /* per picture: Y + U + V */
    uint16_t *pics;             /* pinned: batch pictures */
    uint8_t *bgr;               /* all frames so far (host, pageable, grown on demand) */
    uint16_t *cnt;
    uint8_t *map;
    uint32_t *sum;
    size_t frames, cap;
    size_t mbs, bgr_bytes;
    double gpu_s, copy_s;
    int failed;
    int dry;                    /* JM_INPROC_DRY=1: no GPU at all (timing of the decoder inside this process) */
} H;

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static void hook_flush(void)
{
    double t0;
    int rc;
    if (!H.filled || H.failed) return;
    if (H.frames + (size_t)H.filled > H.cap) {
        H.cap = (H.frames + (size_t)H.filled) * 2;
        H.bgr = (uint8_t *)realloc(H.bgr, H.cap * H.bgr_bytes);
        H.cnt = (uint16_t *)realloc(H.cnt, H.cap * H.mbs * sizeof(uint16_t));
        H.map = (uint8_t *)realloc(H.map, H.cap * H.mbs);
        H.sum = (uint32_t *)realloc(H.sum, H.cap * 2 * sizeof(uint32_t));
        if (!H.bgr || !H.cnt || !H.map || !H.sum) { H.failed = 1; return; }
    }
    if (H.dry) { H.frames += (size_t)H.filled; H.filled = 0; return; }
    t0 = now_s();
    rc = lv_step_host_pictures(H.ctx, H.pics, H.size_x, H.size_y, H.crop_left, H.crop_top, 0, 1, H.filled,
                               H.bgr + H.frames * H.bgr_bytes, NULL, H.cnt + H.frames * H.mbs, H.map + H.frames * H.mbs,
                               H.sum + H.frames * 2);
    H.gpu_s += now_s() - t0;
    if (rc != LV_OK) {
        fprintf(stderr, "jm_inproc: lv_step_host_pictures: %s (%s)\n", lv_strerror(rc), lv_last_error());
        H.failed = 1;
        return;
    }
    H.frames += (size_t)H.filled;
    H.filled = 0;
}

/* Replaces the reference's write_out_picture (lib/JM/output.c:362) at link time. */
void write_out_picture(StorablePicture *p, int p_out_unused)
{
    static const int SubWidthC[4] = {1, 2, 2, 1}, SubHeightC[4] = {1, 2, 1, 1};      /* output.c:364-365 */
    int crop_left = 0, crop_right = 0, crop_top = 0, crop_bottom = 0;
    size_t ny, nc;
    uint16_t *dst;
    double t0;
    (void)p_out_unused;
    if (p->non_existing || H.failed) return;                                          /* output.c:372 */
    if (p->chroma_format_idc != YUV420 || sizeof(imgpel) != sizeof(uint16_t)) {
        fprintf(stderr, "jm_inproc: only 4:2:0 pictures of 16-bit imgpel are handled\n");
        H.failed = 1;
        return;
    }
    if (p->frame_cropping_flag) {                                                     /* output.c:375-380 */
        crop_left = SubWidthC[p->chroma_format_idc] * p->frame_cropping_rect_left_offset;
        crop_right = SubWidthC[p->chroma_format_idc] * p->frame_cropping_rect_right_offset;
        crop_top = SubHeightC[p->chroma_format_idc] * (2 - p->frame_mbs_only_flag) * p->frame_cropping_rect_top_offset;
        crop_bottom = SubHeightC[p->chroma_format_idc] * (2 - p->frame_mbs_only_flag) * p->frame_cropping_rect_bottom_offset;
    }
    ny = (size_t)p->size_x * p->size_y;
    nc = (size_t)p->size_x_cr * p->size_y_cr;
    if (!H.pics) {                                                                    /* first picture: fix the geometry */
        void *pin = NULL;
        H.size_x = p->size_x; H.size_y = p->size_y; H.crop_left = crop_left; H.crop_top = crop_top;
        H.samples = ny + 2 * nc;
        if (p->size_x - crop_left - crop_right != H.width || p->size_y - crop_top - crop_bottom != H.height) {
            fprintf(stderr, "jm_inproc: the stream is %dx%d after cropping, not %dx%d\n", p->size_x - crop_left - crop_right,
                    p->size_y - crop_top - crop_bottom, H.width, H.height);
            H.failed = 1;
            return;
        }
        if (H.dry) pin = malloc((size_t)H.batch * H.samples * sizeof(uint16_t));
        else if (lv_host_alloc(&pin, (size_t)H.batch * H.samples * sizeof(uint16_t)) != LV_OK) { H.failed = 1; return; }
        H.pics = (uint16_t *)pin;
    }
    if (p->size_x != H.size_x || p->size_y != H.size_y || crop_left != H.crop_left || crop_top != H.crop_top) {
        fprintf(stderr, "jm_inproc: picture geometry changed inside the stream\n");
        H.failed = 1;
        return;
    }
    /* three block copies: every plane is one contiguous allocation behind its row pointers (memalloc.c:26-39) */
    t0 = now_s();
    dst = H.pics + (size_t)H.filled * H.samples;
    memcpy(dst, p->imgY[0], ny * sizeof(imgpel));
    memcpy(dst + ny, p->imgUV[0][0], nc * sizeof(imgpel));                            /* U = imgUV[0], output.c:427-436 */
    memcpy(dst + ny + nc, p->imgUV[1][0], nc * sizeof(imgpel));                       /* V = imgUV[1], output.c:438-446 */
    H.copy_s += now_s() - t0;
    if (++H.filled == H.batch) hook_flush();
}

static int dump(const char *prefix, const char *ext, const void *data, size_t bytes)
{
    char name[1024];
    FILE *f;
    snprintf(name, sizeof name, "%s.%s", prefix, ext);
    f = fopen(name, "wb");
    if (!f) { perror(name); return 1; }
    if (bytes && fwrite(data, 1, bytes, f) != bytes) { fclose(f); return 1; }
    return fclose(f) != 0;
}

static void *idle_thread(void *a) { (void)a; for (;;) pause(); return NULL; }

int main(int argc, char **argv)
{
    double t0, t_total;
    int rc;
    if (argc < 6) {
        fprintf(stderr, "usage: %s stream.264 width height batch out_prefix\n", argv[0]);
        return 2;
    }
    H.width = atoi(argv[2]); H.height = atoi(argv[3]); H.batch = atoi(argv[4]);
    if (H.batch < 1) H.batch = 1;
    H.dry = getenv("JM_INPROC_DRY") != NULL;
    if (getenv("JM_INPROC_THREAD")) { pthread_t t; pthread_create(&t, NULL, idle_thread, NULL); }
    rc = H.dry ? LV_OK : lv_create(0, H.width, H.height, 1, 20, 0, &H.ctx);
    if (rc != LV_OK) {
        fprintf(stderr, "jm_inproc: lv_create: %s (%s)\n", lv_strerror(rc), lv_last_error());
        return 3;
    }
    H.mbs = H.dry ? 1 : lv_frame_mbs(H.ctx);
    H.bgr_bytes = H.dry ? 1 : lv_frame_bytes_bgr(H.ctx);
    rc = jmd_open(argv[1], "(in process)", -1, H.width, H.height);
    if (rc) return rc;
    /* The decoder reads its stream one fgetc() at a time (feof + fgetc per byte, lib/JM/annexb.c:65-142: 400 million calls for 64 I_PCM pictures of 1080p).
     * glibc skips the FILE lock while a process has a single thread; the CUDA runtime starts threads, and from then on every
     * call locks: the decoder ran at 7.4 pictures/s in this process against 16 on its own.  Only this thread reads the
     * stream, so take the lock out of the calls (the reference's source is untouched). */
    __fsetlocking(bits, FSETLOCKING_BYCALLER);      /* GetAnnexbNALU (image.c:2263) */
    __fsetlocking(bitss, FSETLOCKING_BYCALLER);     /* partial_GetAnnexbNALU (image.c:2626) */
    t0 = now_s();
    jmd_decode_all();
    hook_flush();                                   /* the last, partial batch */
    t_total = now_s() - t0;
    if (H.failed) return 6;
    if (H.dry) {
        printf("{\"frames\": %zu, \"total_s\": %.6f, \"plane_copy_s\": %.6f, \"dry\": true}\n", H.frames, t_total, H.copy_s);
        return 0;
    }
    if (dump(argv[5], "bgr", H.bgr, H.frames * H.bgr_bytes) || dump(argv[5], "cnt", H.cnt, H.frames * H.mbs * 2) ||
        dump(argv[5], "map", H.map, H.frames * H.mbs) || dump(argv[5], "sum", H.sum, H.frames * 8))
        return 7;
    printf("{\"frames\": %zu, \"picture\": [%d, %d], \"crop_left_top\": [%d, %d], \"imgpel_bytes\": %zu, \"batch\": %d, "
           "\"total_s\": %.6f, \"gpu_stage_s\": %.6f, \"plane_copy_s\": %.6f, \"gpu_launches\": %llu}\n",
           H.frames, H.size_x, H.size_y, H.crop_left, H.crop_top, sizeof(imgpel), H.batch, t_total, H.gpu_s, H.copy_s,
           (unsigned long long)lv_kernel_launches(H.ctx));
    lv_destroy(H.ctx);
    return 0;
}
