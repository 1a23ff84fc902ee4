/*
 * jm_driver.h -- TEST / BENCH INFRASTRUCTURE.  The decoder driver shared by jm_main.c (frames to a file / pipe) and
 * tools/jm_hook/jm_inproc.c (pictures handed to the GPU library inside the same process): what the reference's MFC dialog
 * does around the JM decoder (MainDlg.cpp:296-374) and what the commented-out main of lib/JM/ldecod.c:221-307 did.
 * Compiled only where the reference's headers are present (/root/reference/lib/JM).
 */
#ifndef JM_DRIVER_H
#define JM_DRIVER_H

#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "global.h"
#include "annexb.h"
#include "erc_api.h"
#include "fmo.h"
#include "image.h"
#include "mbuffer.h"
#include "memalloc.h"
#include "output.h"

extern FILE *bitss;                 /* lib/JM/annexb.c:28: the stream handle GetAnnexbNALU reads */
extern objectBuffer_t *erc_object_list;
extern ercVariables_t *erc_errorVar;
extern ColocatedParams *Co_located;
extern int tot_time;

static unsigned char ***jmd_alloc3(int h, int w, int c)
{
    unsigned char ***p = (unsigned char ***)calloc((size_t)h, sizeof(unsigned char **));
    unsigned char **rows = (unsigned char **)calloc((size_t)h * w, sizeof(unsigned char *));
    unsigned char *data = (unsigned char *)calloc((size_t)h * w * c, 1);
    int i, j;
    for (i = 0; i < h; i++) {
        p[i] = rows + (size_t)i * w;
        for (j = 0; j < w; j++) p[i][j] = data + ((size_t)i * w + j) * c;
    }
    return p;
}

/* Everything up to the decode loop.  out_fd: where write_out_picture would write (1 = stdout), or -1 when the pictures
 * are intercepted.  Returns 0 or a small error code. */
static int jmd_open(const char *stream, const char *outname, int out_fd, int w, int h)
{
    int mbs;
    h = (h + 15) / 16 * 16;                       /* coded size: whole macroblocks */
    w = (w + 15) / 16 * 16;
    mbs = (w / 16) * (h / 16);

    input = (struct inp_par *)calloc(1, sizeof(struct inp_par));
    snr = (struct snr_par *)calloc(1, sizeof(struct snr_par));
    img = (struct img_par *)calloc(1, sizeof(struct img_par));
    if (!input || !snr || !img) return 3;

    /* MainDlg.cpp:313-321 */
    strcpy(input->infile, stream);
    strcpy(input->outfile, outname);
    strcpy(input->reffile, "/nonexistent_ref.yuv");
    input->FileFormat = PAR_OF_ANNEXB;
    input->ref_offset = 0;
    input->poc_scale = 2;                          /* bin/decoder.cfg line 7 */
    input->write_uv = 1;
    img->conceal_mode = input->conceal_mode = 0;   /* decoder.cfg: error concealment off */
    img->ref_poc_gap = input->ref_poc_gap = 2;
    img->poc_gap = input->poc_gap = 2;

    /* the analysis globals of lib/JM/global.h:146-238 that decode paths touch; everything else stays 0 = off */
    iwidth = w;
    iheight = h;
    curr_frm = jmd_alloc3(h, w, 3);
    prev_frm = jmd_alloc3(h, w, 3);
    curr_res = jmd_alloc3(h, w, 3);
    MbType = (int *)calloc((size_t)mbs, sizeof(int));
    full_decode_flag = 1;

    p_out = out_fd;
    p_ref = -1;

    init_old_slice();
    OpenBitstreamFile(input->infile);              /* `bits`: read_new_slice calls ftell() on it (image.c:2260) */
    bitss = fopen(input->infile, "rb");            /* MainDlg.cpp:354-355: the handle GetAnnexbNALU reads */
    if (!bitss) {
        perror(input->infile);
        return 5;
    }

    /* lib/JM/ldecod.c:243-270 (the commented-out main) */
    malloc_slice(input, img);
    init(img);
    dec_picture = NULL;
    dpb.init_done = 0;
    g_nFrame = 0;
    init_out_buffer();
    img->idr_psnr_number = input->ref_offset;
    img->psnr_number = 0;
    img->number = 0;
    img->type = I_SLICE;
    img->dec_ref_pic_marking_buffer = NULL;
    Bframe_ctr = snr->frame_ctr = 0;
    tot_time = 0;
    return 0;
}

static void jmd_decode_all(void)
{
    while (decode_one_frame(img, input, snr) != EOS)
        ;
    flush_dpb();
#ifdef PAIR_FIELDS_IN_OUTPUT
    flush_pending_output(p_out);
#endif
    fclose(bitss);
}

#endif /* JM_DRIVER_H */
