// shim_demo.cpp -- the reference's call site, verbatim in shape (MainDlg.cpp:937-946), linked against
// liblivevideo.so through include/lv_convert_shim.hpp.  Build:  make -C tools/harness
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "lv_convert_shim.hpp"

int main(int argc, char **argv)
{
    const int width = argc > 1 ? atoi(argv[1]) : 352, height = argc > 2 ? atoi(argv[2]) : 288;
    const int YSize = height * width, UVSize = YSize / 4;
    std::vector<unsigned char> srcY(YSize, 128), srcU(UVSize, 128), srcV(UVSize, 128), RGBbuf(3 * YSize);
    if (argc > 3) {                                    // optional .yuv file, frame 0
        FILE *f = fopen(argv[3], "rb");
        if (!f || fread(srcY.data(), 1, YSize, f) != (size_t)YSize || fread(srcU.data(), 1, UVSize, f) != (size_t)UVSize ||
            fread(srcV.data(), 1, UVSize, f) != (size_t)UVSize) {
            fprintf(stderr, "cannot read %s\n", argv[3]);
            return 1;
        }
        fclose(f);
    }
    ColorSpaceConversions pCon;
    pCon.YV12_to_RGB24(srcY.data(), srcU.data(), srcV.data(), RGBbuf.data(), width, height);
    printf("B,G,R of the first DIB pixel: %d,%d,%d (expect 130,130,130 for mid grey)\n", RGBbuf[0], RGBbuf[1], RGBbuf[2]);
    std::vector<unsigned char> mask(YSize), map(((width + 15) / 16) * ((height + 15) / 16));
    std::vector<unsigned short> cnt(map.size());
    std::vector<unsigned char> prev(srcY);
    prev[0] = 0;                                        // one pixel differs by 128 > tol
    FrameDiff_MB(srcY.data(), prev.data(), width, height, 20, 0, mask.data(), cnt.data(), map.data());
    printf("changed pixels in macroblock 0: %d, block map: %d\n", cnt[0], map[0]);

    // CMainDlg::DisplayPic (MainDlg.cpp:937-977): five pictures, the dialog's own pointer arrays, one pipelined call
    unsigned char *srcYs[5], *srcUs[5], *srcVs[5], *RGBbufs[5];
    std::vector<std::vector<unsigned char>> own;
    for (int i = 0; i < 5; i++) {
        own.emplace_back(YSize, (unsigned char)(16 + 40 * i)); srcYs[i] = own.back().data();
        own.emplace_back(UVSize, 128); srcUs[i] = own.back().data();
        own.emplace_back(UVSize, 128); srcVs[i] = own.back().data();
        own.emplace_back(3 * YSize); RGBbufs[i] = own.back().data();
    }
    const int rc = pCon.YV12_to_RGB24(srcYs, srcUs, srcVs, RGBbufs, 5, width, height);
    printf("five pictures in one call: rc %d, first blue bytes %d %d %d %d %d (expect 0 46 93 139 186)\n", rc, RGBbufs[0][0],
           RGBbufs[1][0], RGBbufs[2][0], RGBbufs[3][0], RGBbufs[4][0]);
    return rc;
}
