/*
 * harness32.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Freestanding i386 driver around the reference's ORIGINAL object code:
 * /root/reference/cscc.lib -> ./Temp/Convert.obj (MSVC 1997, thiscall), which
 * implements class ColorSpaceConversions declared in /root/reference/Convert.h:19-31.
 * No reference source is copied: the object is extracted, converted to ELF32 and
 * linked from where it lies by oracle/ref32/build_ref.py; the output binary goes
 * to oracle/_ref/oracle32 (git-ignored).
 *
 * There is no 32-bit libc in the image, so this file uses raw int $0x80 syscalls.
 *
 * usage:  oracle32 MODE W H [REPS]  < input  > output
 *   MODE 0  YV12_to_RGB24  in: N x I420 frames (Y,U,V planes)   out: N x 3*W*H BGR bottom-up
 *   MODE 1  RGB24_to_YV12  in: N x 3*W*H                         out: N x 1.5*W*H
 *   MODE 2  YVU9_to_YV12 / 3 YUY2_to_YV12 / 4 YV12_to_YVU9 / 5 YV12_to_YUY2 (in/out sizes below)
 *   REPS>0: read ONE frame, run the call REPS times, write the output once and
 *           print "s=<sec> ns=<nsec>" (elapsed) on stderr (CPU-baseline timing of the real object code).
 */
typedef unsigned char u8;
typedef unsigned int u32;

int fltused_sym __asm__("__fltused") = 0x9875;

/* MSVC CRT _ftol: truncate st(0) toward zero, pop, result in edx:eax. */
__asm__(".text\n.globl __ftol\n__ftol:\n"
        " sub $12,%esp\n fnstcw (%esp)\n movw (%esp),%ax\n orb $0x0c,%ah\n"
        " movw %ax,2(%esp)\n fldcw 2(%esp)\n fistpll 4(%esp)\n fldcw (%esp)\n"
        " movl 4(%esp),%eax\n movl 8(%esp),%edx\n add $12,%esp\n ret\n");

#define THISCALL __attribute__((thiscall))
extern THISCALL void *csc_ctor(void *self);
extern THISCALL void csc_yv12_rgb24(void *self, u8 *y, u8 *u, u8 *v, u8 *dst, int w, int h);
extern THISCALL void csc_rgb24_yv12(void *self, u8 *in, u8 *out, int w, int h);
extern THISCALL void csc_yvu9_yv12(void *self, u8 *in, u8 *out, int w, int h);
extern THISCALL void csc_yuy2_yv12(void *self, u8 *in, u8 *out, int w, int h);
extern THISCALL void csc_yv12_yvu9(void *self, u8 *in, u8 *out, int w, int h);
extern THISCALL void csc_yv12_yuy2(void *self, u8 *in, u8 *out, int w, int h);

static int sys3(int nr, int a, int b, int c)
{
    int r;
    __asm__ volatile("int $0x80" : "=a"(r) : "0"(nr), "b"(a), "c"(b), "d"(c) : "memory");
    return r;
}
static int rd_full(u8 *p, u32 n)
{
    u32 got = 0;
    while (got < n) {
        int r = sys3(3, 0, (int)(p + got), (int)(n - got));
        if (r <= 0) return (int)got;
        got += (u32)r;
    }
    return (int)got;
}
static void wr_full(int fd, const u8 *p, u32 n)
{
    u32 done = 0;
    while (done < n) {
        int r = sys3(4, fd, (int)(p + done), (int)(n - done));
        if (r <= 0) sys3(1, 3, 0, 0);
        done += (u32)r;
    }
}
static int parse_int(const char *s)
{
    int v = 0;
    while (*s >= '0' && *s <= '9') v = v * 10 + (*s++ - '0');
    return v;
}
struct ts32 { int sec; int nsec; };
static void now(struct ts32 *t) { sys3(265, 1 /*CLOCK_MONOTONIC*/, (int)t, 0); }
static int put_u32(char *dst, u32 v)
{
    char tmp[12]; int k = 0, n = 0;
    if (v == 0) tmp[k++] = '0';
    while (v) { tmp[k++] = (char)('0' + (int)(v % 10u)); v /= 10u; }
    while (k) dst[n++] = tmp[--k];
    return n;
}

#define MAXPX (3840 * 2160 + 4096)
static u8 inbuf[MAXPX * 3 + 64];
static u8 outbuf[MAXPX * 3 + 64];
static u8 self[16];

static void run(int mode, int w, int h)
{
    u32 px = (u32)w * (u32)h;
    switch (mode) {
    case 0: csc_yv12_rgb24(self, inbuf, inbuf + px, inbuf + px + px / 4, outbuf, w, h); break;
    case 1: csc_rgb24_yv12(self, inbuf, outbuf, w, h); break;
    case 2: csc_yvu9_yv12(self, inbuf, outbuf, w, h); break;
    case 3: csc_yuy2_yv12(self, inbuf, outbuf, w, h); break;
    case 4: csc_yv12_yvu9(self, inbuf, outbuf, w, h); break;
    default: csc_yv12_yuy2(self, inbuf, outbuf, w, h); break;
    }
}

int cmain(int argc, char **argv)
{
    if (argc < 4) return 2;
    int mode = parse_int(argv[1]), w = parse_int(argv[2]), h = parse_int(argv[3]);
    int reps = argc > 4 ? parse_int(argv[4]) : 0;
    u32 px = (u32)w * (u32)h;
    if (w <= 0 || h <= 0 || px > MAXPX - 4096 || mode < 0 || mode > 5) return 2;
    /* bytes in / out per frame: I420 = 1.5 B/px, RGB24 = 3, YVU9 = 1.125, YUY2 = 2 */
    static const u32 in_num[6]  = {12, 24,  9, 16, 12, 12};
    static const u32 out_num[6] = {24, 12, 12, 12,  9, 16};
    u32 in_sz = px / 8 * in_num[mode], out_sz = px / 8 * out_num[mode];
    csc_ctor(self);
    if (reps > 0) {
        if (rd_full(inbuf, in_sz) != (int)in_sz) return 4;
        struct ts32 t0, t1;
        now(&t0);
        for (int i = 0; i < reps; i++) run(mode, w, h);
        now(&t1);
        wr_full(1, outbuf, out_sz);
        int ds = t1.sec - t0.sec, dn = t1.nsec - t0.nsec;
        if (dn < 0) { ds -= 1; dn += 1000000000; }
        char msg[48]; int n = 0;
        msg[n++] = 's'; msg[n++] = '='; n += put_u32(msg + n, (u32)ds);
        msg[n++] = ' '; msg[n++] = 'n'; msg[n++] = 's'; msg[n++] = '='; n += put_u32(msg + n, (u32)dn);
        msg[n++] = '\n';
        wr_full(2, (u8 *)msg, (u32)n);
        return 0;
    }
    for (;;) {
        int got = rd_full(inbuf, in_sz);
        if (got == 0) break;
        if (got != (int)in_sz) return 4;
        run(mode, w, h);
        wr_full(1, outbuf, out_sz);
    }
    return 0;
}

__asm__(".text\n.globl _start\n_start:\n"
        " mov (%esp),%eax\n lea 4(%esp),%edx\n and $-16,%esp\n sub $8,%esp\n"
        " push %edx\n push %eax\n call cmain\n"
        " mov %eax,%ebx\n mov $1,%eax\n int $0x80\n");
