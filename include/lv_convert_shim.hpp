// lv_convert_shim.hpp -- source-compatible replacement for the reference's `Convert.h` + `cscc.lib`.
//
// The MFC views hold a `ColorSpaceConversions pCon;` by value (MainDlg.h:68, BkdextDlg.h:40) and call
// `pCon.YV12_to_RGB24(srcY, srcU, srcV, RGBbuf, width, height)` (MainDlg.cpp:946,953,960,967,974;
// BkdextDlg.cpp:106).  Including this header instead of Convert.h and linking liblivevideo instead of
// cscc.lib keeps those call sites unchanged: the class has the same name, the same public methods with
// the same parameter lists (Convert.h:19-31), no data members and a trivial constructor (the original's
// constructor only fills process-global tables, Convert.obj .text+0x000..0x1cf; the CUDA path needs none).
//
// Only YV12_to_RGB24 is ever called by the application; the other five converters forward to their C entries too
// (bit-exact with the object code, but simple kernels: they are not on the hot path).
#ifndef LV_CONVERT_SHIM_HPP
#define LV_CONVERT_SHIM_HPP

#include "livevideo.h"

class ColorSpaceConversions {
public:
    ColorSpaceConversions() {}

    // planar 4:2:0 (src0 = Y, src1 = U/Cb, src2 = V/Cr) -> 24-bit bottom-up B,G,R, bit-exact with cscc.lib
    void YV12_to_RGB24(unsigned char *src0, unsigned char *src1, unsigned char *src2, unsigned char *dst_ori,
                       int width, int height)
    {
        ::YV12_to_RGB24(src0, src1, src2, dst_ori, width, height);
    }

    // DisplayPic's five conversions (MainDlg.cpp:946-974) in ONE call: pCon.YV12_to_RGB24(srcY, srcU, srcV, RGBbuf, 5,
    // width, height) with the dialog's own pointer arrays; uploads, kernels and downloads of the pictures overlap.
    // Returns LV_OK or an LV_E_* code.
    int YV12_to_RGB24(unsigned char *const *src0, unsigned char *const *src1, unsigned char *const *src2,
                      unsigned char *const *dst_ori, int n, int width, int height)
    {
        return ::lv_convert_host_multi(src0, src1, src2, dst_ori, n, width, height);
    }

    // the five converters the application never calls; same object-code semantics, same signatures
    void RGB24_to_YV12(unsigned char *in, unsigned char *out, int w, int h) { ::RGB24_to_YV12(in, out, w, h); }
    void YVU9_to_YV12(unsigned char *in, unsigned char *out, int w, int h) { ::YVU9_to_YV12(in, out, w, h); }
    void YUY2_to_YV12(unsigned char *in, unsigned char *out, int w, int h) { ::YUY2_to_YV12(in, out, w, h); }
    void YV12_to_YVU9(unsigned char *in, unsigned char *out, int w, int h) { ::YV12_to_YVU9(in, out, w, h); }
    void YV12_to_YUY2(unsigned char *in, unsigned char *out, int w, int h) { ::YV12_to_YUY2(in, out, w, h); }
};

#endif  // LV_CONVERT_SHIM_HPP
