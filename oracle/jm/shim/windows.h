/* minimal stand-in for <windows.h> so the reference's JM sources compile on Linux */
#ifndef LV_SHIM_WINDOWS_H
#define LV_SHIM_WINDOWS_H
#include <stdint.h>
#include <time.h>
typedef unsigned char BYTE;
typedef union { struct { uint32_t LowPart; int32_t HighPart; } u; long long QuadPart; } LARGE_INTEGER;
static inline int QueryPerformanceCounter(LARGE_INTEGER *t) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); t->QuadPart = (long long)ts.tv_sec * 1000000000LL + ts.tv_nsec; return 1; }
static inline int QueryPerformanceFrequency(LARGE_INTEGER *f) { f->QuadPart = 1000000000LL; return 1; }
#ifndef max
#define max(a, b) (((a) > (b)) ? (a) : (b))
#endif
#ifndef min
#define min(a, b) (((a) < (b)) ? (a) : (b))
#endif
#endif
