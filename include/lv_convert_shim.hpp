// lv_convert_shim.hpp -- source-compatible replacement for the reference's `Convert.h` + `cscc.lib`.
//
// The MFC views hold a `ColorSpaceConversions pCon;` by value (MainDlg.h:68, BkdextDlg.h:40) and call
// `pCon.YV12_to_RGB24(srcY, srcU, srcV, RGBbuf, width, height)` (MainDlg.cpp:946,953,960,967,974;
// BkdextDlg.cpp:106).  Including this header instead of Convert.h and linking liblivevideo instead of
// cscc.lib keeps those call sites unchanged: the class has the same name, the same public methods with
// the same parameter lists (Convert.h:19-31), no data members and a trivial constructor (the original's
// constructor only fills process-global tables, Convert.obj .text+0x000..0x1cf; the CUDA path needs none).
//
// Only YV12_to_RGB24 is ever called by the application; the other five converters are declared so that
// existing code compiles, and report that they are not part of the B200 path if somebody calls them.
#ifndef LV_CONVERT_SHIM_HPP
#define LV_CONVERT_SHIM_HPP

#include <cstdio>
#include <cstdlib>

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

    void RGB24_to_YV12(unsigned char *, unsigned char *, int, int) { not_on_path("RGB24_to_YV12"); }
    void YVU9_to_YV12(unsigned char *, unsigned char *, int, int) { not_on_path("YVU9_to_YV12"); }
    void YUY2_to_YV12(unsigned char *, unsigned char *, int, int) { not_on_path("YUY2_to_YV12"); }
    void YV12_to_YVU9(unsigned char *, unsigned char *, int, int) { not_on_path("YV12_to_YVU9"); }
    void YV12_to_YUY2(unsigned char *, unsigned char *, int, int) { not_on_path("YV12_to_YUY2"); }

private:
    static void not_on_path(const char *name)
    {
        std::fprintf(stderr, "livevideo-b200: ColorSpaceConversions::%s is not part of the accelerated path "
                             "(the application never calls it)\n", name);
        std::abort();
    }
};

#endif  // LV_CONVERT_SHIM_HPP
