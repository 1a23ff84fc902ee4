/*
 * lv_oracle.c -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY (see lv_oracle.h for scope and citations).
 * Never linked into, imported by or called from the product path.
 */
#include "lv_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * Tables of ColorSpaceConversions::ColorSpaceConversions()  (Convert.obj .text+0x110..0x1c1).
 * `>>` on negative ints is the x86 `sar` of the original: arithmetic (floor) shift.
 * ---------------------------------------------------------------------------------------- */
static int32_t crv_tab[256];    /* (104597*(i-128))>>15   .text+0x14a/0x152 */
static int32_t cbu_tab[256];    /* (132201*(i-128))>>15   .text+0x15b/0x163 */
static int32_t cgu_tab[256];    /*  25675*(i-128)          unshifted, .text+0x12d/0x145 */
static int32_t cgv_tab[256];    /*  53279*(i-128)          unshifted, .text+0x133/0x13f */
static int32_t tab_76309[256];  /* ( 76309*(i- 16))>>15   .text+0x16c/0x174 */
static uint8_t clp_base[2048];  /* clp = clp_base + 0x300, x in [-768,1280)  .text+0x180..0x1c1 */
static int tables_ready;

static inline int32_t sar(int32_t v, int s)
{
    /* floor shift, independent of the compiler's treatment of signed >> */
    return v >= 0 ? (v >> s) : -(int32_t)(((uint32_t)(-(v + 1)) >> s) + 1);
}

/* The hot loops use the compiler's signed >> directly; lvo_init_tables() verifies once that it is the
 * arithmetic (floor) shift of the original's `sar`, as gcc/clang guarantee on x86-64. */
#define SAR15(v) ((int32_t)(v) >> 15)
#if defined(__GNUC__) && defined(__x86_64__) && !defined(LVO_NO_CLONES)
#define LVO_HOT __attribute__((target_clones("avx2", "default")))
#else
#define LVO_HOT
#endif

void lvo_init_tables(void)
{
    if (tables_ready) return;
    {
        volatile int32_t m1 = -1, m3 = -98305;
        if ((m1 >> 1) != -1 || (m3 >> 15) != -4) abort();   /* signed >> must be arithmetic */
    }
    for (int i = 0; i < 256; i++) {
        crv_tab[i] = sar(104597 * (i - 128), 15);
        cbu_tab[i] = sar(132201 * (i - 128), 15);
        cgu_tab[i] = 25675 * (i - 128);
        cgv_tab[i] = 53279 * (i - 128);
        tab_76309[i] = sar(76309 * (i - 16), 15);
    }
    for (int x = -768; x < 1280; x++) {
        uint8_t c;
        if (x < 0) c = 0;                 /* .text+0x196 test/jl -> dl(=0) */
        else if (x > 510) c = 255;        /* .text+0x19a cmp eax,0x1fe; jle */
        else c = (uint8_t)(x >> 1);       /* .text+0x1ab sar ebx,1 */
        clp_base[x + 0x300] = c;
    }
    tables_ready = 1;
}

#define CLP(x) (clp_base[(x) + 0x300])

void lvo_yuv_to_bgr_px(int y, int u, int v, uint8_t bgr[3])
{
    lvo_init_tables();
    int32_t t = tab_76309[y];
    int32_t g = sar(cgu_tab[u] + cgv_tab[v], 15);
    bgr[0] = CLP(t + cbu_tab[u]);
    bgr[1] = CLP(t - g);
    bgr[2] = CLP(t + crv_tab[v]);
}

void lvo_textbook_px(int y, int u, int v, uint8_t bgr[3])
{
    int c = 76309 * (y - 16), d = u - 128, e = v - 128;
    int r = sar(c + 104597 * e, 16), g = sar(c - 25675 * d - 53279 * e, 16), b = sar(c + 132201 * d, 16);
    bgr[0] = (uint8_t)(b < 0 ? 0 : b > 255 ? 255 : b);
    bgr[1] = (uint8_t)(g < 0 ? 0 : g > 255 ? 255 : g);
    bgr[2] = (uint8_t)(r < 0 ? 0 : r > 255 ? 255 : r);
}

/*
 * ColorSpaceConversions::YV12_to_RGB24, Convert.obj .text+0x350..0x73f.
 * Walks 4x2 pixel blocks; src row 0 lands in the LAST dst row (bottom-up DIB), bytes B,G,R.
 * src1 is U (Cb), src2 is V (Cr) -- the callers pass srcU, srcV (MainDlg.cpp:946).
 * Loop counts are (w+3)>>2 and (h+1)>>1, so the original needs w%4==0 and h%2==0 to stay in
 * bounds; the restatement keeps the same traversal and therefore the same precondition.
 */
LVO_HOT void lvo_yv12_to_rgb24(const uint8_t *src_y, const uint8_t *src_u, const uint8_t *src_v,
                               uint8_t *dst, int width, int height)
{
    lvo_init_tables();
    const uint8_t *py1 = src_y, *py2 = src_y + width;
    uint8_t *d1 = dst + (size_t)3 * width * (height - 1);   /* .text+0x350..0x39f; size_t: 2^30-pixel pictures pass 2^31 bytes */
    uint8_t *d2 = d1 - (size_t)3 * width;
    const int back = (9 * width) & ~3;              /* .text+0x3a5..0x3b8 */
    for (int j = (height + 1) >> 1; j > 0; j--) {
        for (int i = (width + 3) >> 2; i > 0; i--) {
            int32_t cb[2], cr[2], cg[2];
            for (int k = 0; k < 2; k++) {
                int u = src_u[k], v = src_v[k];
                cb[k] = cbu_tab[u];
                cr[k] = crv_tab[v];
                cg[k] = SAR15(cgu_tab[u] + cgv_tab[v]);     /* add, then ONE sar: +0x412, +0x45c */
            }
            for (int p = 0; p < 4; p++) {
                int k = p >> 1;
                int32_t t1 = tab_76309[py1[p]], t2 = tab_76309[py2[p]];
                d1[3 * p + 0] = CLP(t1 + cb[k]);
                d1[3 * p + 1] = CLP(t1 - cg[k]);
                d1[3 * p + 2] = CLP(t1 + cr[k]);
                d2[3 * p + 0] = CLP(t2 + cb[k]);
                d2[3 * p + 1] = CLP(t2 - cg[k]);
                d2[3 * p + 2] = CLP(t2 + cr[k]);
            }
            d1 += 12; d2 += 12; py1 += 4; py2 += 4; src_u += 2; src_v += 2;   /* +0x6a6..0x6dd */
        }
        d1 -= back; d2 -= back;          /* +3w forward in the row, -9w back => net -6w */
        py1 += width; py2 += width;
    }
}

/* ------------------------------------------------------------------------------------------ */
int lvo_is_nearly_zero(int a, int tol)
{
    /* #define IsNearlyZero(a,tol) (((a)>=0)&&((a)<=(tol)))||(((a)<0)&&((a)>=-(tol)))   image.h:22 */
    return ((a >= 0) && (a <= tol)) || ((a < 0) && (a >= -tol));
}

/* !IsNearlyZero(a, tol) with the macro body of image.h:22 written out, branch-free so the row loop vectorises */
static inline int changed_px(int a, int tol)
{
    return !((((a) >= 0) & ((a) <= (tol))) | (((a) < 0) & ((a) >= -(tol))));
}

LVO_HOT uint32_t lvo_frame_diff_mb(const uint8_t *cur_y, const uint8_t *ref_y, int width, int height,
                                   int tol, int mb_thresh, uint8_t *mask, uint16_t *mb_count, uint8_t *mb_map)
{
    const int mb_w = (width + 15) / 16, mb_h = (height + 15) / 16;
    uint32_t total = 0;
    /* Same result as scanning each 16x16 block (image.c:4978-4981, 5143-5150: k = mby*mbwidth + mbx,
     * py in [16*mby, 16*mby+16), px in [16*mbx, 16*mbx+16)), traversed row by row so every pixel is
     * touched once in memory order; per pixel image.c:4795-4803 / MainDlg.cpp:806 (uchar promotes to int). */
    uint32_t *acc = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)mb_w);
    for (int mby = 0; mby < mb_h; mby++) {
        memset(acc, 0, sizeof(uint32_t) * (size_t)mb_w);
        for (int py = mby * 16; py < mby * 16 + 16 && py < height; py++) {
            const uint8_t *c = cur_y + (size_t)py * width, *r = ref_y + (size_t)py * width;
            uint8_t *mrow = mask ? mask + (size_t)py * width : NULL;
            for (int mbx = 0; mbx < mb_w; mbx++) {
                const int x0 = mbx * 16, x1 = x0 + 16 < width ? x0 + 16 : width;
                uint32_t cnt = 0;
                if (mrow) {
                    for (int px = x0; px < x1; px++) {
                        int m = changed_px((int)c[px] - (int)r[px], tol);
                        mrow[px] = (uint8_t)m;
                        cnt += (uint32_t)m;
                    }
                } else {
                    for (int px = x0; px < x1; px++) cnt += (uint32_t)changed_px((int)c[px] - (int)r[px], tol);
                }
                acc[mbx] += cnt;
            }
        }
        for (int mbx = 0; mbx < mb_w; mbx++) {
            const int k = mby * mb_w + mbx;
            if (mb_count) mb_count[k] = (uint16_t)acc[mbx];
            if (mb_map) mb_map[k] = (uint8_t)((int)acc[mbx] > mb_thresh);
            total += acc[mbx];
        }
    }
    free(acc);
    return total;
}

void lvo_pack_mask16(const uint8_t *mask, int width, int height, uint16_t *bits)
{
    int units = width / 16;
    for (int y = 0; y < height; y++)
        for (int u = 0; u < units; u++) {
            uint16_t b = 0;
            for (int i = 0; i < 16; i++) b |= (uint16_t)((mask[y * width + u * 16 + i] & 1) << i);
            bits[y * units + u] = b;
        }
}

/* ------------------------------------------------------------------------------------------ */
typedef struct {
    const uint8_t *i420; uint8_t *prev_luma; int n_streams, n_frames, width, height, tol, mb_thresh;
    uint8_t *bgr, *mask; uint16_t *mb_count; uint8_t *mb_map; uint32_t *summary;
    int next; pthread_mutex_t mu;
} batch_job;

static void one_frame(batch_job *j, int s, int t)
{
    const size_t px = (size_t)j->width * j->height, fsz = px + px / 2;
    const int mbs = ((j->width + 15) / 16) * ((j->height + 15) / 16);
    const size_t f = (size_t)s * j->n_frames + t;
    const uint8_t *y = j->i420 + f * fsz, *u = y + px, *v = u + px / 4;
    const uint8_t *ref = t == 0 ? j->prev_luma + (size_t)s * px : y - fsz;
    if (j->bgr) lvo_yv12_to_rgb24(y, u, v, j->bgr + f * 3 * px, j->width, j->height);
    uint16_t *cnt = j->mb_count ? j->mb_count + f * mbs : NULL;
    uint8_t *map = j->mb_map ? j->mb_map + f * mbs : NULL;
    uint16_t *tmp = NULL;
    if (!cnt && j->summary) cnt = tmp = (uint16_t *)malloc(sizeof(uint16_t) * mbs);
    uint32_t changed = lvo_frame_diff_mb(y, ref, j->width, j->height, j->tol, j->mb_thresh,
                                         j->mask ? j->mask + f * px : NULL, cnt, map);
    if (j->summary) {
        uint32_t nmb = 0;
        for (int k = 0; k < mbs; k++) nmb += (uint32_t)((int)cnt[k] > j->mb_thresh);
        j->summary[2 * f] = changed;
        j->summary[2 * f + 1] = nmb;
    }
    free(tmp);
}

static void *batch_worker(void *arg)
{
    batch_job *j = (batch_job *)arg;
    const int total = j->n_streams * j->n_frames;
    for (;;) {
        pthread_mutex_lock(&j->mu);
        int i = j->next++;
        pthread_mutex_unlock(&j->mu);
        if (i >= total) break;
        one_frame(j, i / j->n_frames, i % j->n_frames);   /* one frame per task */
    }
    return NULL;
}

void lvo_convert_diff_batch(const uint8_t *i420, uint8_t *prev_luma, int n_streams, int n_frames,
                            int width, int height, int tol, int mb_thresh,
                            uint8_t *bgr, uint8_t *mask, uint16_t *mb_count, uint8_t *mb_map,
                            uint32_t *summary, int n_threads)
{
    lvo_init_tables();
    batch_job j = {i420, prev_luma, n_streams, n_frames, width, height, tol, mb_thresh,
                   bgr, mask, mb_count, mb_map, summary, 0, PTHREAD_MUTEX_INITIALIZER};
    if (n_threads <= 1) {
        batch_worker(&j);
    } else {
        pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * n_threads);
        for (int i = 0; i < n_threads; i++) pthread_create(&th[i], NULL, batch_worker, &j);
        for (int i = 0; i < n_threads; i++) pthread_join(th[i], NULL);
        free(th);
    }
    /* prev_frm <- curr_frm (image.c:1553-1561): every frame only read prev_luma or earlier inputs,
     * so the state update happens after the whole batch. */
    const size_t px = (size_t)width * height, fsz = px + px / 2;
    for (int s = 0; s < n_streams && n_frames > 0; s++)
        memcpy(prev_luma + (size_t)s * px, i420 + ((size_t)s * n_frames + n_frames - 1) * fsz, px);
}

/* ------------------------------------------------------------------------------------------ */
void lvo_write_bks(const uint8_t *cur, const uint8_t *bkd, int width, int height, int tol, uint8_t *out)
{
    /* MainDlg.cpp:804-824: three flat loops over Y, U, V of one I420 frame */
    const int ysz = width * height, uvsz = ysz / 4, voff = ysz + uvsz;
    for (int j = 0; j < ysz; j++)
        out[j] = lvo_is_nearly_zero((int)cur[j] - (int)bkd[j], tol) ? 16 : cur[j];
    for (int j = 0; j < uvsz; j++)
        out[j + ysz] = lvo_is_nearly_zero((int)cur[j + ysz] - (int)bkd[j + ysz], tol) ? 128 : cur[j + ysz];
    for (int j = 0; j < uvsz; j++)
        out[j + voff] = lvo_is_nearly_zero((int)cur[j + voff] - (int)bkd[j + voff], tol) ? 128 : cur[j + voff];
}

/* ------------------------------------------------------------------------------------------
 * The other five converters (object code only; see lv_oracle.h).
 * ---------------------------------------------------------------------------------------- */
static uint32_t r2yuv[256], g2yuv[256], b2yuv[256];
static int rgb_tables_ready;

static void init_rgb_tables(void)
{
    /* ctor, Convert.obj .text+0x025..0x10a: three packed tables, fields of 10 bits: Y | Cb<<10 | Cr<<20.  Every
     * entry is __ftol(coef * i) (truncation toward zero) of the doubles in .rdata; negative contributions are
     * kept as their low 8 bits, the field sums are taken modulo 256 by the byte stores of RGB24_to_YV12. */
    if (rgb_tables_ready) return;
    for (int i = 0; i < 256; i++) {
        const uint32_t half = (uint32_t)(int)(0.5 * i);
        r2yuv[i] = (half << 20) | (((uint32_t)(int)(-0.1687 * i) & 0xff) << 10) | (uint32_t)(int)(0.299 * i);
        g2yuv[i] = (((uint32_t)(int)(-0.4187 * i) & 0xff) << 20) | (((uint32_t)(int)(-0.3313 * i) & 0xff) << 10) |
                   (uint32_t)(int)(0.587 * i);
        b2yuv[i] = (((uint32_t)(int)(-0.0813 * i) & 0xff) << 20) | (half << 10) | (uint32_t)(int)(0.114 * i);
    }
    rgb_tables_ready = 1;
}

void lvo_rgb24_to_yv12(const uint8_t *in, uint8_t *out, int w, int h)
{
    /* .text+0x1d0..0x344: 2x2 blocks, starting at the LAST DIB row (the top image row); luma for all four
     * pixels, chroma from the top-left pixel only, XOR 0x80 */
    init_rgb_tables();
    const int px = w * h;
    uint8_t *py = out, *pu = out + px, *pv = out + px + px / 4;
    const uint8_t *row = in + 3 * w * (h - 1);
    for (int j = h >> 1; j > 0; j--) {
        const uint8_t *p = row;
        for (int i = (w + 1) >> 1; i > 0; i--) {
            const uint8_t *q = p - 3 * w;                       /* the image row below = previous DIB row */
            uint32_t s = r2yuv[p[0]] + g2yuv[p[1]] + b2yuv[p[2]];   /* bytes are R,G,B here (+0x23e..0x25a) */
            py[0] = (uint8_t)s;
            *pu++ = (uint8_t)((s >> 10) ^ 0x80);
            *pv++ = (uint8_t)((s >> 20) ^ 0x80);
            py[w] = (uint8_t)(r2yuv[q[0]] + g2yuv[q[1]] + b2yuv[q[2]]);
            py[1] = (uint8_t)(r2yuv[p[3]] + g2yuv[p[4]] + b2yuv[p[5]]);
            py[w + 1] = (uint8_t)(r2yuv[q[3]] + g2yuv[q[4]] + b2yuv[q[5]]);
            py += 2;
            p += 6;
        }
        py += w;
        row -= 6 * w;
    }
}

void lvo_yvu9_to_yv12(const uint8_t *in, uint8_t *out, int w, int h)
{
    /* .text+0x740..0x8a2: every plane is read bottom-up; chroma (w/4 x h/4) is replicated 2x2 */
    const int px = w * h, cw4 = w / 4, ch4 = h / 4, cw2 = w / 2;
    for (int y = 0; y < h; y++) memcpy(out + (size_t)y * w, in + (size_t)(h - 1 - y) * w, (size_t)w);
    for (int pl = 0; pl < 2; pl++) {
        const uint8_t *src = in + px + pl * (h * cw4 / 4);
        uint8_t *dst = out + px + pl * (h * cw2 / 2);
        for (int y = 0; y < ch4; y++) {
            const uint8_t *s = src + (size_t)(ch4 - 1 - y) * cw4;
            uint8_t *d = dst + (size_t)(2 * y) * cw2;
            for (int x = 0; x < cw4; x++) d[2 * x] = d[2 * x + 1] = s[x];
            memcpy(d + cw2, d, (size_t)cw2);
        }
    }
}

void lvo_yuy2_to_yv12(const uint8_t *in, uint8_t *out, int w, int h)
{
    /* .text+0x8b0..0x956: Y0 U Y1 V; byte 3 feeds the FIRST chroma plane, byte 1 the second; both rows of a pair
     * write the same chroma row, so the second (odd) row's chroma survives */
    const int px = w * h, cw2 = w / 2;
    uint8_t *py = out, *p1 = out + px, *p2 = out + px + px / 4;
    for (int j = (h + 1) >> 1; j > 0; j--) {
        for (int r = 0; r < 2; r++) {
            for (int i = 0; i < (w + 1) >> 1; i++) {
                *py++ = in[0]; p2[i] = in[1]; *py++ = in[2]; p1[i] = in[3];
                in += 4;
            }
        }
        p1 += cw2; p2 += cw2;
    }
}

void lvo_yv12_to_yvu9(const uint8_t *in, uint8_t *out, int w, int h)
{
    /* .text+0x960..0xa39: luma copied; each chroma plane keeps every second sample of every second row */
    const int px = w * h, cw4 = w / 4, ch4 = h / 4, cw2 = w / 2;
    memcpy(out, in, (size_t)px);
    const uint8_t *s = in + px;
    uint8_t *d = out + px;
    for (int pl = 0; pl < 2; pl++)
        for (int y = 0; y < ch4; y++) {
            for (int x = 0; x < cw4; x++) *d++ = s[2 * x];
            s += 2 * cw4 + cw2;
        }
}

void lvo_yv12_to_yuy2(const uint8_t *in, uint8_t *out, int w, int h)
{
    /* .text+0xa40..0xae6: inverse of YUY2_to_YV12's layout; a chroma row serves both rows of a pair */
    const int px = w * h, cw2 = w / 2;
    const uint8_t *py = in, *p1 = in + px, *p2 = in + px + px / 4;
    for (int j = (h + 1) >> 1; j > 0; j--) {
        for (int r = 0; r < 2; r++)
            for (int i = 0; i < (w + 1) >> 1; i++) {
                *out++ = *py++; *out++ = p2[i]; *out++ = *py++; *out++ = p1[i];
            }
        p1 += cw2; p2 += cw2;
    }
}

void lvo_object_box(const uint8_t *mask, int width, int x0, int x1, int y0, int y1, int box[4])
{
    /* image.c:4818-4838 */
    int max_obj_x = x0 - 1, min_obj_x = x1 + 1, max_obj_y = y0 - 1, min_obj_y = y1 + 1;
    for (int m = y0; m <= y1; m++)
        for (int n = x0; n <= x1; n++)
            if (mask[m * width + n] == 1) {
                if (max_obj_x < n) max_obj_x = n;
                if (min_obj_x > n) min_obj_x = n;
                if (max_obj_y < m) max_obj_y = m;
                if (min_obj_y > m) min_obj_y = m;
            }
    box[0] = min_obj_x; box[1] = max_obj_x; box[2] = min_obj_y; box[3] = max_obj_y;
}

/* The object bookkeeping either side of the box in motion_segment's I-frame branch.
 * image.h:23-24: #define ClipX(x) (x<0?0:(x>=iwidth?iwidth-1:x)), ClipY likewise with iheight. */
static int clip_to(int x, int n) { return x < 0 ? 0 : (x >= n ? n - 1 : x); }

void lvo_object_window(const int obj[4], int width, int height, int hor_margin, int ver_margin, int win[4])
{
    /* image.c:4786-4789; obj = {PosX, PosY, Width, Height}; win = {min_x, max_x, min_y, max_y} */
    win[0] = clip_to(obj[0] - obj[2] / 2 - hor_margin, width);
    win[1] = clip_to(obj[0] + obj[2] / 2 + hor_margin, width);
    win[2] = clip_to(obj[1] - obj[3] / 2 - ver_margin, height);
    win[3] = clip_to(obj[1] + obj[3] / 2 + ver_margin, height);
}

void lvo_object_update(const int box[4], int obj[4])
{
    /* image.c:4840-4848; box = {min_obj_x, max_obj_x, min_obj_y, max_obj_y} */
    const int min_obj_x = box[0], max_obj_x = box[1], min_obj_y = box[2], max_obj_y = box[3];
    if ((min_obj_x < max_obj_x) && (min_obj_y < max_obj_y)) {
        obj[0] = (int)(min_obj_x + (max_obj_x - min_obj_x) / 2);
        obj[1] = (int)(min_obj_y + (max_obj_y - min_obj_y) / 2);
        obj[2] = (int)(max_obj_x - min_obj_x);
        obj[3] = (int)(max_obj_y - min_obj_y);
    }
}

/* ------------------------------------------------------------------------------------------
 * Generators; all arithmetic uint32 wrap-around.
 * ---------------------------------------------------------------------------------------- */
uint32_t lvo_lowbias32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

static uint32_t key0(uint32_t seed, uint32_t stream) { return lvo_lowbias32(seed + 0x9E3779B9U * (stream + 1)); }

static void fill_uniform(uint32_t k0, uint32_t frame, uint32_t plane, uint8_t *dst, size_t n)
{
    uint32_t k = lvo_lowbias32(k0 ^ (3 * frame + plane));
    for (size_t i = 0; i < n; i++) dst[i] = (uint8_t)(lvo_lowbias32(k + (uint32_t)i) >> 24);
}

void lvo_gen_uniform(uint32_t seed, uint32_t stream, uint32_t frame, int width, int height, uint8_t *i420)
{
    const size_t px = (size_t)width * height;
    const uint32_t k0 = key0(seed, stream);
    fill_uniform(k0, frame, 0, i420, px);
    fill_uniform(k0, frame, 1, i420 + px, px / 4);
    fill_uniform(k0, frame, 2, i420 + px + px / 4, px / 4);
}

void lvo_gen_camera(uint32_t seed, uint32_t stream, uint32_t frame, int width, int height, uint8_t *i420)
{
    static const int amp[4] = {0, 6, 12, 24};
    const size_t px = (size_t)width * height;
    const uint32_t k0 = key0(seed, stream);
    const uint32_t k = lvo_lowbias32(k0 ^ (3 * frame + 0));
    const uint32_t kbg = lvo_lowbias32(k0 ^ 0xB5297A4DU ^ 0u);
    const int mb_w = (width + 15) / 16;
    const int bw = width / 8, bh = height / 8;
    const int x0 = (int)((8 * frame) % (uint32_t)(width - bw)), y0 = (int)((4 * frame) % (uint32_t)(height - bh));
    for (int y = 0; y < height; y++)
        for (int x = 0; x < width; x++) {
            uint32_t i = (uint32_t)(y * width + x);
            int bg = 32 + (int)((lvo_lowbias32(kbg + i) >> 24) * 3 / 4);
            int val;
            if (x >= x0 && x < x0 + bw && y >= y0 && y < y0 + bh) {
                val = 255 - bg;
            } else {
                int a = amp[lvo_lowbias32(k0 ^ 0x68E31DA4U ^ (uint32_t)((y / 16) * mb_w + x / 16)) & 3];
                uint32_t n = lvo_lowbias32(k + i) >> 8;
                val = bg + (int)(n % (uint32_t)(2 * a + 1)) - a;
                if (val < 0) val = 0;
                if (val > 255) val = 255;
            }
            i420[i] = (uint8_t)val;
        }
    fill_uniform(k0, frame, 1, i420 + px, px / 4);
    fill_uniform(k0, frame, 2, i420 + px + px / 4, px / 4);
}

uint32_t lvo_crc32(const uint8_t *p, uint64_t n)
{
    static uint32_t tab[256];
    static int ready;
    if (!ready) {
        for (uint32_t i = 0; i < 256; i++) {
            uint32_t c = i;
            for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320U ^ (c >> 1) : c >> 1;
            tab[i] = c;
        }
        ready = 1;
    }
    uint32_t c = 0xFFFFFFFFU;
    for (uint64_t i = 0; i < n; i++) c = tab[(c ^ p[i]) & 0xFF] ^ (c >> 8);
    return c ^ 0xFFFFFFFFU;
}

/* ---- decoder picture -> planar 8-bit I420 ----------------------------------------------------------------------- */
/* img2buf, lib/JM/output.c:87-178.  On a little-endian host with sizeof(imgpel) == 2 > symbol_size_in_bytes == 1 the
 * function takes the last branch (output.c:161-176): size = 1 and
 *     memcpy(buf + (j - crop_left + (i - crop_top) * twidth), &imgX[i][j], 1)
 * i.e. the first byte in memory of the short = its low byte. */
void lvo_img2buf(const uint16_t *plane, int size_x, int size_y, int crop_left, int crop_right, int crop_top,
                 int crop_bottom, uint8_t *buf)
{
    const int twidth = size_x - crop_left - crop_right;
    for (int i = crop_top; i < size_y - crop_bottom; i++)
        for (int j = crop_left; j < size_x - crop_right; j++)
            buf[(j - crop_left) + (size_t)(i - crop_top) * twidth] = (uint8_t)(plane[(size_t)i * size_x + j] & 0xffu);
}

/* write_out_picture, lib/JM/output.c:362-453, for chroma_format_idc == 1 (4:2:0) and YUV output: the luma crop is
 * SubWidthC/SubHeightC (= 2) times the SPS offsets (output.c:375-380), the chroma crop the offsets themselves
 * (output.c:421-425); planes leave in the order Y (427-428), imgUV[0] (437-438), imgUV[1] (446-447). */
void lvo_picture16_to_i420(const uint16_t *pic16, int size_x, int size_y, int crop_left, int crop_right, int crop_top,
                           int crop_bottom, uint8_t *i420)
{
    const int w = size_x - crop_left - crop_right, h = size_y - crop_top - crop_bottom;
    const int cx = size_x / 2, cy = size_y / 2;
    const uint16_t *u16 = pic16 + (size_t)size_x * size_y, *v16 = u16 + (size_t)cx * cy;
    lvo_img2buf(pic16, size_x, size_y, crop_left, crop_right, crop_top, crop_bottom, i420);
    lvo_img2buf(u16, cx, cy, crop_left / 2, crop_right / 2, crop_top / 2, crop_bottom / 2, i420 + (size_t)w * h);
    lvo_img2buf(v16, cx, cy, crop_left / 2, crop_right / 2, crop_top / 2, crop_bottom / 2,
                i420 + (size_t)w * h + (size_t)(w / 2) * (h / 2));
}
