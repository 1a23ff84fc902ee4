/*
 * jm_main.c -- TEST / BENCH INFRASTRUCTURE.  A `main()` for the reference's bundled JM 10.2 decoder
 * (/root/reference/lib/JM, compiled UNMODIFIED from where it lies by oracle/jm/build_jm.py; the
 * reference's own main() is commented out, lib/JM/ldecod.c:221-307, because the MFC dialog drives the
 * decoder, MainDlg.cpp:296-374).  It reproduces that driver on Linux so the decoder can act as the
 * host-side producer of planar I420 frames for BASELINE.json configs[4]:
 *
 *     jm_ldecod  stream.264  out.yuv|-  width height
 *
 *     jm_ldecod  --img2buf size_x size_y crop_left crop_right crop_top crop_bottom  < plane.u16  > plane.u8
 *         runs the reference's OWN img2buf (lib/JM/output.c:87-178) with symbol_size_in_bytes = 1 on one imgpel plane
 *         read from stdin: the known-answer source that pins the oracle's restatement of the narrowing + cropping
 *
 *     jm_ldecod  --nearlyzero  > table.u8
 *         evaluates the reference's OWN macro IsNearlyZero (lib/JM/image.h:22, as compiled here) for every difference
 *         a in [-255, 255] and every tol in [0, 255]: 511 x 256 bytes of 0/1, row = a + 255.  Pins the oracle's
 *         lvo_is_nearly_zero, i.e. the per-pixel rule "changed <=> !IsNearlyZero(cur - ref, tol)".
 *
 * Frames leave the decoder through write_out_picture (lib/JM/output.c:362-453): Y, U, V, 8 bit, cropped.
 * "-" writes them to stdout (a pipe to the GPU harness).  width/height size the author's analysis
 * buffers that his hooks in the macroblock decoder write unconditionally (curr_frm, curr_res:
 * lib/JM/macroblock.c:3916-3936, 6025-6069; allocated by CMainDlg::AllocBuf in the application).
 */
#include "jm_driver.h"

void img2buf(imgpel **imgX, unsigned char *buf, int size_x, int size_y, int symbol_size_in_bytes, int crop_left,
             int crop_right, int crop_top, int crop_bottom);          /* lib/JM/output.c:87 */

static int img2buf_mode(char **a)
{
    const int sx = atoi(a[0]), sy = atoi(a[1]), cl = atoi(a[2]), cr = atoi(a[3]), ct = atoi(a[4]), cb = atoi(a[5]);
    const size_t n = (size_t)sx * sy, on = (size_t)(sx - cl - cr) * (sy - ct - cb);
    imgpel *data = (imgpel *)malloc(n * sizeof(imgpel));
    imgpel **rows = (imgpel **)malloc((size_t)sy * sizeof(imgpel *));
    unsigned char *out = (unsigned char *)malloc(n);
    size_t got = 0, put = 0;
    int i;
    if (sx <= 0 || sy <= 0 || !data || !rows || !out) return 3;
    while (got < n * sizeof(imgpel)) {
        ssize_t r = read(0, (char *)data + got, n * sizeof(imgpel) - got);
        if (r <= 0) return 4;
        got += (size_t)r;
    }
    for (i = 0; i < sy; i++) rows[i] = data + (size_t)i * sx;          /* the layout of get_mem2Dpel, memalloc.c:26-39 */
    img2buf(rows, out, sx, sy, 1, cl, cr, ct, cb);
    while (put < on) {
        ssize_t r = write(1, out + put, on - put);
        if (r <= 0) return 5;
        put += (size_t)r;
    }
    return 0;
}

static int nearlyzero_mode(void)
{
    static unsigned char t[511 * 256];
    int a, tol;
    size_t put = 0;
    for (a = -255; a <= 255; a++)
        for (tol = 0; tol <= 255; tol++) t[(a + 255) * 256 + tol] = IsNearlyZero(a, tol) ? 1 : 0;   /* image.h:22 */
    while (put < sizeof t) {
        ssize_t r = write(1, t + put, sizeof t - put);
        if (r <= 0) return 5;
        put += (size_t)r;
    }
    return 0;
}

int main(int argc, char **argv)
{
    int fd, rc;
    if (argc == 8 && strcmp(argv[1], "--img2buf") == 0) return img2buf_mode(argv + 2);
    if (argc == 2 && strcmp(argv[1], "--nearlyzero") == 0) return nearlyzero_mode();
    if (argc < 5) {
        fprintf(stderr, "usage: %s stream.264 out.yuv|- width height\n", argv[0]);
        return 2;
    }
    if (strcmp(argv[2], "-") == 0) fd = 1;         /* stdout */
    else if ((fd = open(argv[2], O_WRONLY | O_CREAT | O_TRUNC, 0644)) == -1) {
        perror(argv[2]);
        return 4;
    }
    rc = jmd_open(argv[1], argv[2], fd, atoi(argv[3]), atoi(argv[4]));
    if (rc) return rc;
    jmd_decode_all();
    if (fd != 1) close(fd);
    return 0;
}
