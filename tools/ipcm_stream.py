#!/usr/bin/env python3
"""Synthetic H.264 Annex-B writer: every macroblock is I_PCM, so the decoded pictures are exactly the input
frames.  No encoder exists on the build or GPU boxes; this gives the bundled JM decoder (the host-side producer
of BASELINE.json configs[4]) a bitstream whose decoded I420 output is known byte for byte.

Stream: SPS (baseline, POC type 2, frame cropping when the height is not a multiple of 16) + PPS (CAVLC,
deblocking control present) + one IDR slice per frame (slice_type I, deblocking disabled, mb_type 25 = I_PCM).
"""
import numpy as np


class BitWriter:
    def __init__(self):
        self.out = bytearray()
        self.acc = 0
        self.n = 0

    def u(self, bits, val):
        for i in range(bits - 1, -1, -1):
            self.acc = (self.acc << 1) | ((val >> i) & 1)
            self.n += 1
            if self.n == 8:
                self.out.append(self.acc)
                self.acc, self.n = 0, 0

    def ue(self, v):
        v += 1
        nbits = v.bit_length()
        self.u(nbits - 1, 0)
        self.u(nbits, v)

    def se(self, v):
        self.ue(2 * v - 1 if v > 0 else -2 * v)

    def align_zero(self):
        while self.n:
            self.u(1, 0)

    def bytes(self, b):
        assert self.n == 0
        self.out += b

    def trailing(self):
        self.u(1, 1)
        self.align_zero()


def _escape(rbsp: bytes) -> bytes:
    """emulation prevention: 00 00 0x (x <= 3) -> 00 00 03 0x"""
    a = np.frombuffer(rbsp, dtype=np.uint8)
    if a.size < 3:
        return rbsp
    # candidate positions: two zeros followed by a byte <= 3; insertion resets the zero run, so walk the candidates
    cand = np.flatnonzero((a[:-2] == 0) & (a[1:-1] == 0) & (a[2:] <= 3))
    if cand.size == 0:
        return rbsp
    out = bytearray()
    last = 0
    zeros_consumed_until = -1
    for c in cand:
        # the two zeros at c, c+1 must not overlap a pair already consumed by a previous insertion
        if c < zeros_consumed_until:
            continue
        out += rbsp[last:c + 2]
        out.append(3)
        last = c + 2
        zeros_consumed_until = c + 2
    out += rbsp[last:]
    return bytes(out)


def _nal(ref_idc, typ, rbsp):
    return b"\x00\x00\x00\x01" + bytes([(ref_idc << 5) | typ]) + _escape(rbsp)


def sps(width, height, level_idc=40):
    mb_w, mb_h = (width + 15) // 16, (height + 15) // 16
    b = BitWriter()
    b.u(8, 66)            # profile_idc baseline
    b.u(8, 0)             # constraint flags + reserved
    b.u(8, level_idc)
    b.ue(0)               # seq_parameter_set_id
    b.ue(0)               # log2_max_frame_num_minus4
    b.ue(2)               # pic_order_cnt_type 2: output order == decoding order
    b.ue(1)               # num_ref_frames
    b.u(1, 0)             # gaps_in_frame_num_value_allowed_flag
    b.ue(mb_w - 1)
    b.ue(mb_h - 1)
    b.u(1, 1)             # frame_mbs_only_flag
    b.u(1, 1)             # direct_8x8_inference_flag
    crop_b, crop_r = (mb_h * 16 - height) // 2, (mb_w * 16 - width) // 2
    if crop_b or crop_r:
        b.u(1, 1)
        b.ue(0); b.ue(crop_r); b.ue(0); b.ue(crop_b)
    else:
        b.u(1, 0)
    b.u(1, 0)             # vui_parameters_present_flag
    b.trailing()
    return _nal(3, 7, bytes(b.out))


def pps():
    b = BitWriter()
    b.ue(0); b.ue(0)      # pps id, sps id
    b.u(1, 0)             # entropy_coding_mode_flag: CAVLC
    b.u(1, 0)             # pic_order_present_flag
    b.ue(0)               # num_slice_groups_minus1
    b.ue(0); b.ue(0)      # num_ref_idx_l0/l1_active_minus1
    b.u(1, 0)             # weighted_pred_flag
    b.u(2, 0)             # weighted_bipred_idc
    b.se(0); b.se(0); b.se(0)   # pic_init_qp/qs, chroma_qp_index_offset
    b.u(1, 1)             # deblocking_filter_control_present_flag
    b.u(1, 0)             # constrained_intra_pred_flag
    b.u(1, 0)             # redundant_pic_cnt_present_flag
    b.trailing()
    return _nal(3, 8, bytes(b.out))


def idr_slice(i420, width, height, idr_pic_id):
    mb_w, mb_h = (width + 15) // 16, (height + 15) // 16
    cw, ch = mb_w * 16, mb_h * 16
    px = width * height
    y = np.zeros((ch, cw), np.uint8)
    u = np.full((ch // 2, cw // 2), 128, np.uint8)
    v = np.full((ch // 2, cw // 2), 128, np.uint8)
    y[:height, :width] = i420[:px].reshape(height, width)
    u[:height // 2, :width // 2] = i420[px:px + px // 4].reshape(height // 2, width // 2)
    v[:height // 2, :width // 2] = i420[px + px // 4:].reshape(height // 2, width // 2)
    # macroblock-major sample order: 256 luma, 64 Cb, 64 Cr per macroblock
    yb = y.reshape(mb_h, 16, mb_w, 16).transpose(0, 2, 1, 3).reshape(mb_h * mb_w, 256)
    ub = u.reshape(mb_h, 8, mb_w, 8).transpose(0, 2, 1, 3).reshape(mb_h * mb_w, 64)
    vb = v.reshape(mb_h, 8, mb_w, 8).transpose(0, 2, 1, 3).reshape(mb_h * mb_w, 64)
    b = BitWriter()
    b.ue(0)               # first_mb_in_slice
    b.ue(7)               # slice_type: I (all slices of the picture)
    b.ue(0)               # pic_parameter_set_id
    b.u(4, 0)             # frame_num (IDR)
    b.ue(idr_pic_id)
    b.u(1, 0); b.u(1, 0)  # no_output_of_prior_pics_flag, long_term_reference_flag
    b.se(0)               # slice_qp_delta
    b.ue(1)               # disable_deblocking_filter_idc = 1
    # slice header done; macroblock layer.  After the header the writer is generally not byte aligned, but every
    # macroblock is ue(25) = 9 bits followed by pcm_alignment_zero_bits and 384 aligned bytes, so only the first
    # macroblock's prefix needs the bit writer; the rest is assembled with numpy.
    b.ue(25)
    b.align_zero()
    head = bytes(b.out)
    mbs = np.concatenate([yb, ub, vb], axis=1)                        # [n_mb, 384]
    # ue(25) = 0000 1 1010 (9 bits) then 7 alignment zeros = bytes 0x0D 0x00
    prefix = np.tile(np.array([0x0D, 0x00], np.uint8), (mbs.shape[0], 1))
    body = np.concatenate([prefix, mbs], axis=1).reshape(-1)[2:]      # first macroblock's prefix is in `head`
    rbsp = head + body.tobytes() + b"\x80"                            # rbsp_slice_trailing_bits
    return _nal(3, 5, rbsp)


def write_stream(path, frames, width, height):
    """frames: iterable of I420 uint8 arrays."""
    with open(path, "wb") as f:
        f.write(sps(width, height))
        f.write(pps())
        for k, fr in enumerate(frames):
            f.write(idr_slice(np.asarray(fr, dtype=np.uint8), width, height, k & 1))


if __name__ == "__main__":
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
    import lvoracle as orc
    w, h, n, out = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    write_stream(out, (orc.gen("camera", 1, 0, t, w, h) for t in range(n)), w, h)
    print(f"wrote {out}: {n} frames {w}x{h}")
