// lv_kernels.h -- internal launch interface between the C ABI (lv_capi.cu) and the kernels.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace lv {

struct Geom {
    int w, h;          // pixels
    int units;         // w / 16 (16-pixel column units)
    int mb_w, mb_h;    // ceil(w/16), ceil(h/16)   (lib/JM/image.c:4677-4679 uses floor on MB-aligned sizes)
    size_t px;         // w*h
    size_t frame;      // 1.5*w*h, bytes of one I420 frame
};

inline Geom make_geom(int w, int h)
{
    Geom g;
    g.w = w; g.h = h; g.units = w / 16;
    g.mb_w = (w + 15) / 16; g.mb_h = (h + 15) / 16;
    g.px = (size_t)w * h; g.frame = g.px + g.px / 2;
    return g;
}

// Decoder pictures as the input of the fused step: 16-bit samples (imgpel = unsigned short, lib/JM/global.h:61), planes
// Y (size_x x size_y), U, V (half size each way) back to back, `samples` apart; the frame is the crop window.
struct PicLayout {
    const uint16_t *pic16 = nullptr;
    int size_x = 0, size_y = 0, crop_left = 0, crop_top = 0;
    size_t samples = 0;
};
bool pic_fusable(const Geom &g, const PicLayout &p);

// One launch of the fused step over n_streams*n_frames frames (stream-major).
struct StepArgs {
    Geom g;
    const uint8_t *i420;      // input frames
    const uint8_t *state_in;  // previous luma planes, n_streams x px (slot of stream s = s)
    uint8_t *state_out;       // luma of the last frame of each stream is written here (may be NULL)
    int n_streams, n_frames;
    int static_ref;           // 1: every frame of a stream is differenced against state_in (never against frame t-1)
    int tol, mb_thresh;
    uint8_t *bgr;             // any output may be NULL
    uint16_t *mask_bits;
    uint8_t *mask_bytes;
    uint16_t *mb_count;
    uint8_t *mb_map;
    uint32_t *summary;        // 2 per frame; must be zeroed before the launch
    // object boxes accumulated by the fused kernel (optional; boxes must be initialised: launch_step_init)
    const int *windows;       // n_obj x {x0, x1, y0, y1} per frame, inclusive, 16-byte aligned
    int *boxes;               // n_obj x {min_x, max_x, min_y, max_y} per frame, 16-byte aligned
    int n_obj;                // 1 .. 16
    int box_windowed;         // 1: plain fused step + a pass over the windows only (k_box_windows) instead of the fused box
    PicLayout pic;            // pic.pic16 != NULL: the frames are decoder pictures (i420 unused, bgr mandatory, no boxes)
};

// difference of explicit (cur, ref) plane pairs: frame i of cur_y against frame i of ref_y
struct DiffArgs {
    Geom g;
    const uint8_t *cur_y, *ref_y;
    int n_frames;
    int tol, mb_thresh;
    uint16_t *mask_bits;
    uint8_t *mask_bytes;
    uint16_t *mb_count;
    uint8_t *mb_map;
    uint32_t *summary;
};

// width % 16 == 0, height % 2 == 0.  Return the number of kernels launched (>0) or a negative
// cudaError_t value (-(int)err).
int launch_step(const StepArgs &a, cudaStream_t st);
int launch_step_init(uint32_t *summary, size_t n_frames, const int *windows, int *boxes, int n_obj, cudaStream_t st);
int launch_convert(const Geom &g, const uint8_t *i420, int n_frames, uint8_t *bgr, cudaStream_t st);
int launch_diff(const DiffArgs &a, cudaStream_t st);

// lv_zerocopy.cu: conversion of up to 8 pictures whose planes are separate allocations the device can address directly
// (host memory pinned and mapped in place): y/dst 16-byte aligned, u/v 8-byte aligned.
struct PlaneTable {
    static const int kMax = 8;
    const uint8_t *y[kMax], *u[kMax], *v[kMax];
    uint8_t *dst[kMax];
};
int launch_convert_planes(const Geom &g, const PlaneTable &t, int n, bool staged_stores, cudaStream_t st);

int kernel_variant();
void set_kernel_variant(int v);

// The TMA pipeline (lv_tma.cu).  Return 0 when the geometry or a pointer alignment is not eligible
// (the caller then uses the tiled kernel), >0 = kernels launched, <0 = -(cudaError_t).
bool tma_supported(const Geom &g);
int launch_step_tma(const StepArgs &a, cudaStream_t st);
int launch_convert_tma(const Geom &g, const uint8_t *i420, int n_frames, uint8_t *bgr, cudaStream_t st);
int launch_diff_tma(const DiffArgs &a, cudaStream_t st);

// lv_aux.cu: background subtraction (WriteBks) and object box
int launch_write_bks(const Geom &g, const uint8_t *cur, const uint8_t *bkd, int n_frames, int tol, uint8_t *out,
                     uint16_t *obj_bits, cudaStream_t st);
int launch_narrow(const Geom &g, const uint16_t *pic16, int n_frames, int size_x, int size_y, int crop_left, int crop_top,
                  uint8_t *i420, cudaStream_t st);
int launch_object_box(int w, int h, const uint8_t *mask, bool bits, int n_frames, int x0, int x1, int y0, int y1, int *d_boxes,
                      cudaStream_t st);

// the object bookkeeping either side of the box (lib/JM/image.c:4786-4789 and 4840-4848); objects = {PosX, PosY, Width, Height}
int launch_object_windows(int w, int h, const int *objects, int n, int hor_margin, int ver_margin, int *windows, cudaStream_t st);
int launch_object_update(const int *boxes, int n, int *objects, cudaStream_t st);

// lv_formats.cu: the other five cscc.lib converters (kind 1 RGB24->YV12, 2 YVU9->YV12, 3 YUY2->YV12, 4 YV12->YVU9,
// 5 YV12->YUY2)
size_t format_bytes_in(int kind, int w, int h);
size_t format_bytes_out(int kind, int w, int h);
int launch_format(int kind, int w, int h, const uint8_t *in, uint8_t *out, int n_frames, cudaStream_t st);

// any width % 4 == 0 / height % 2 == 0 (legacy geometry that is not a multiple of 16)
int launch_convert_generic(int w, int h, const uint8_t *y, const uint8_t *u, const uint8_t *v,
                           uint8_t *bgr, cudaStream_t st);
int launch_diff_generic(int w, int h, const uint8_t *cur, const uint8_t *ref, int tol, int mb_thresh,
                        uint8_t *mask_bytes, uint16_t *mb_count, uint8_t *mb_map, cudaStream_t st);

}  // namespace lv
