// xengine.cu -- clXEngine: the X stage of an FX correlator.
//
// Reference: CharToComplex (lib/clXEngine_impl.cc:819-916) expands the int8
// integration buffer to complex float in global memory (4x the bytes), then
// XCorrelate (:708-817) runs one work-item per (channel, baseline), each
// streaming 2*T strided global loads with no reuse.
//
// Here: for every frequency channel f the integration is the real int8 matrix
//     Z_f[(input v, re|im)][t]            (2*NV rows, T columns),
// and all visibilities of the channel come from the Gram matrix G = Z Z^T:
//     Re V(v1,v2) = G[(v1,re),(v2,re)] + G[(v1,im),(v2,im)]
//     Im V(v1,v2) = G[(v1,im),(v2,re)] - G[(v1,re),(v2,im)]          (:729-736)
// computed EXACTLY on the int8 tensor cores (s8 x s8 -> s32).  The input layout
// [t][station][chan][pol][re,im] has time outermost, the tensor cores want time
// innermost, so a CTA stages a tile of FC channels x 32 time steps: coalesced
// 32-bit loads, a 4x4 byte transpose in registers (PRMT), bank-conflict-free
// stores into a per-channel K-major shared-memory image; one warp then owns one
// channel (or a share of its row tiles), loads each 16x32 operand fragment ONCE
// and uses it both as the A operand of its own row tile and as the B operand of
// every row tile at or below it -- rows are ordered [8 inputs re | 8 inputs im]
// so that the re/im combination above happens inside one thread's accumulators.
// Only the lower triangle of 8x8-input blocks is computed.
// HBM traffic: the int8 input once + the visibilities once (71.4 MB
// per integration at 32 stations x 1024 channels x 1024 time steps).
#include "common.cuh"
#include "fft_device.cuh"      // static_for
#include <cmath>
#include <cstdlib>

using namespace clb200;
using clb200::fftdev::static_for;

namespace {

constexpr int XE_WARPS = 16;
constexpr int XE_THREADS = XE_WARPS * 32;
constexpr int XE_TT = 32;             // time steps per stage = one k32 MMA step
constexpr int XE_RSW = 12;            // row stride in words: 32 B data + 16 B pad

__device__ __forceinline__ void mma_s8(int (&d)[4], const int (&a)[4], int b0, int b1)
{
    asm volatile(
        "mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
        "{%0,%1,%2,%3};\n"
        : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct XeParams {
    const int8_t *in;       // [t][station][Fstride][npol][2]
    int32_t *out_i32;       // [F][nbl][npol*npol][2]   (may be null)
    float2 *out_f32;        // same shape (may be null)
    int A, npol, F, Fstride, f_off, T;
    int accumulate;         // out += result
    int split;              // CTAs share channel groups: partial sums meet through int32 atomics
    int nslice;             // > 0: time-sliced decomposition, grid = groups x nslice
    int nbatch;             // integrations back to back in `in` (and matrices in `out`), one grid (TMA kernel, nslice > 0)
    int l2_rows;            // prefetch whole (t, station) rows into L2 ahead of the demand loads
    int aligned;            // rows are 4-byte aligned -> 32-bit loads
    float scale;            // 1/127^2 (IChar) or 1/7^2 (packed 4 bit)
    // fused all-gather (TMA kernel only): the float result also goes to `ngather` full matrices
    // (this rank's and the peers', peer-mapped) at item offset gather_off
    float2 *gather[8];
    int ngather;
    long gather_off;
    // ... the result may instead leave through ONE NVSwitch multicast store per 16 B (gather_mc: multicast address of
    // the full matrix, mapped on every rank), and the completion of this rank's slab is announced to every rank by a
    // release store of `gather_epoch` into flag slot [my rank] of their flag arrays once the LAST CTA has stored
    float2 *gather_mc;
    unsigned *gather_flag[8];       // per destination rank: address of ITS flag word for this source rank (or null)
    unsigned *gather_flag_mc;       // multicast address of the flag word for this source rank (or null)
    unsigned *gather_counter;       // CTAs of this launch that have finished their stores
    int gather_fence_gpu;
    int pdl_nowait;                 // consecutive launches are independent: no griddepcontrol.wait
    unsigned gather_epoch;
};

struct XeGatherSync {
    unsigned *flag[8];          // rank r's flag word for this source rank
    unsigned *flag_mc;          // multicast address of the same word (or null)
    const unsigned *local;      // this rank's flag array
    unsigned epoch;
    int nranks, signal, wait;
};

// CTA-wide barrier of k_xengine_i8: every warp reaches it at the same program location (the warp shares differ only
// inside the MMA / epilogue sections of xe_body), as compute-sanitizer's synccheck requires
__device__ __forceinline__ void xe_cta_sync() { __syncthreads(); }

// which row tiles warp share Q of WPC owns
template <int WPC, int Q>
__host__ __device__ constexpr bool owns(int mi)
{
    int r = mi % (2 * WPC);
    return r == Q || r == 2 * WPC - 1 - Q;
}
template <int MT, int WPC, int Q>
__host__ __device__ constexpr int tile_base(int mi)     // accumulator tiles before row tile mi
{
    int n = 0;
    for (int m = 0; m < mi; m++)
        if (owns<WPC, Q>(m)) n += m + 1;
    return n;
}
template <int MT, int WPC, int Q>
__host__ __device__ constexpr int mi_max()
{
    int mx = -1;
    for (int m = 0; m < MT; m++)
        if (owns<WPC, Q>(m)) mx = m;
    return mx;
}

// MT: row tiles (8 inputs each) per channel; WPC: warps sharing one channel; FC = 16/WPC
template <int MT, int WPC>
__host__ __device__ constexpr int nt_max()                   // most accumulator tiles any warp share owns
{
    int mx = 0;
    for (int q = 0; q < WPC; q++) {
        int n = 0;
        for (int m = 0; m < MT; m++) {
            const int r = m % (2 * WPC);
            if (r == q || r == 2 * WPC - 1 - q) n += m + 1;
        }
        if (n > mx) mx = n;
    }
    return mx;
}

// One body for every warp share: the barriers sit in code all warps run; only the MMA and epilogue sections
// branch on the warp's share q (compile-time tile sets inside each branch).
template <int MT, int WPC, int NPOL>
__device__ __forceinline__ void xe_body(const XeParams &p, uint32_t *zbuf, const int q)
{
    constexpr int FC = XE_WARPS / WPC;
    constexpr int NT = nt_max<MT, WPC>();                    // accumulator tiles (the largest share's)
    constexpr int MMAX = MT - 1;
    constexpr int CSW = MT * 16 * XE_RSW + 2;                // channel stride (words), = 2 mod 32
    constexpr int ZW = FC * CSW;                             // words per stage buffer
    constexpr int NVP = MT * 8;
    constexpr int npol = NPOL;
    constexpr int ASTN = NVP / NPOL;                         // padded station count
    constexpr int RUNW = FC * NPOL / 2;                      // 32-bit words per (t, station) run
    constexpr int NQUAD = 8 * ASTN * RUNW;                   // 4-t quads per stage
    constexpr int QPT = (NQUAD + XE_THREADS - 1) / XE_THREADS;
    static_assert(RUNW >= 1, "a stage row must hold at least one 32-bit word");

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, tig = lane & 3;
    const int chl = warp / WPC;                              // channel within the CTA
    const int nbl = p.A * (p.A + 1) / 2;
    const long rowb = (long)p.Fstride * npol * 2;            // bytes per (t, station)
    const long frameb = rowb * p.A;                          // bytes per t
    const int ngroups = (p.F + FC - 1) / FC;

    // per-thread, stage-invariant description of the 4-t quads it moves:
    // quad e -> (word wi of the run, station s, quad-in-stage q); lanes run over
    // (wi: RUNW values, q low 2 bits) so that both the global loads (32 B runs) and
    // the shared stores (bank = 4*wi' + q) are conflict free
    unsigned qsrc[QPT];      // byte offset from the stage base (32 x frame bytes < 4 GiB, checked on the host)
    int qz[QPT];             // word offset of the (first channel, re) row in a stage buffer
    int qt[QPT];             // first time step of the quad within the stage
    int qwi[QPT];            // word index (channel validity depends on the group)
    bool qok[QPT];
#pragma unroll
    for (int i = 0; i < QPT; i++) {
        const int e = threadIdx.x + i * XE_THREADS;
        const int wi = e % RUNW, r1 = e / RUNW;
        const int qlo = r1 & 3, r2 = r1 >> 2;
        const int s = r2 % ASTN, qhi = r2 / ASTN;
        const int q = qhi * 4 + qlo;
        const int ch_a = (NPOL == 1) ? 2 * wi : wi;
        const int v_a = (NPOL == 1) ? s : 2 * s;
        qsrc[i] = (unsigned)((4 * q) * frameb + s * rowb + wi * 4);
        qz[i] = ch_a * CSW + ((v_a >> 3) * 16 + (v_a & 7)) * XE_RSW + q;
        qt[i] = 4 * q;
        qwi[i] = wi;
        qok[i] = e < NQUAD && s < p.A;
    }
    // second sample of a word: next channel (1 pol) or the Y polarisation of the same channel
    constexpr int ZB = (NPOL == 1) ? CSW : XE_RSW;

    // Stream-K style decomposition: the (channel group, 32-step stage) pairs form one
    // sequence, group-major; CTA c takes an equal contiguous share of it.  A share
    // that does not cover a group's whole integration adds its partial sums into the
    // int32 result with atomics (p.split); otherwise every CTA owns whole groups.
    const int nst = (p.T + XE_TT - 1) / XE_TT;
    int s0, s1;                                   // ngroups * nst < 2^31 (checked on the host)
    if (p.nslice > 0) {
        // time-sliced: CTA = (slice, channel group), groups fastest, one wave.  CTAs that run
        // together read neighbouring 32 B runs of the SAME (t, station) rows at about the same
        // time, which is what lets the DRAM controllers serve them from open pages.
        const int grp = blockIdx.x % ngroups, sl = blockIdx.x / ngroups;
        const int len = (nst + p.nslice - 1) / p.nslice;
        s0 = grp * nst + sl * len;
        s1 = min(grp * nst + nst, s0 + len);
    } else if (p.split) {
        const long total = (long)ngroups * nst;
        s0 = (int)(total * blockIdx.x / gridDim.x);
        s1 = (int)(total * (blockIdx.x + 1) / gridDim.x);
    } else {
        s0 = (int)((long)ngroups * blockIdx.x / gridDim.x) * nst;
        s1 = (int)((long)ngroups * (blockIdx.x + 1) / gridDim.x) * nst;
    }
    if (s0 >= s1) return;

    int acc[NT > 0 ? NT : 1][2][4];
    auto zero_acc = [&]() {
#pragma unroll
        for (int i = 0; i < (NT > 0 ? NT : 1); i++)
#pragma unroll
            for (int c = 0; c < 2; c++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][c][j] = 0;
    };
    zero_acc();

    uint32_t pre[QPT][4];
    auto load_stage = [&](int sg) {
        const int grp = sg / nst, st = sg - grp * nst;
        const int f0 = grp * FC;
        const int8_t *sbase = p.in + ((long)(p.f_off + f0)) * npol * 2 + (long)st * XE_TT * frameb;
        const int trem = p.T - st * XE_TT;           // time steps left from the stage start
        // fast path: whole group in range, no padded stations, 32-bit aligned rows, full stage
        const bool full = p.aligned && (f0 + FC <= p.F) && (p.A == ASTN) && (NQUAD % XE_THREADS == 0);
        if (full && trem >= XE_TT) {
            const unsigned fb = (unsigned)frameb;
#pragma unroll
            for (int i = 0; i < QPT; i++)
#pragma unroll
                for (int k = 0; k < 4; k++)
                    pre[i][k] = __ldg(reinterpret_cast<const unsigned int *>(sbase + (qsrc[i] + k * fb)));
            return;
        }
#pragma unroll
        for (int i = 0; i < QPT; i++) {
            const int chw = (NPOL == 1) ? 2 * qwi[i] : qwi[i];
            const bool ok = qok[i] && (f0 + chw) < p.F;
            const bool ok2 = (NPOL == 1) ? (f0 + chw + 1) < p.F : true;
            const int8_t *src = sbase + qsrc[i];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                uint32_t w = 0;
                if (ok && (qt[i] + k) < trem) {
                    const int8_t *q = src + (long)k * frameb;
                    if (p.aligned && ok2) {
                        w = __ldg(reinterpret_cast<const unsigned int *>(q));
                    } else {
                        w = (uint32_t)(uint8_t)q[0] | ((uint32_t)(uint8_t)q[1] << 8);
                        if (ok2) w |= ((uint32_t)(uint8_t)q[2] << 16) | ((uint32_t)(uint8_t)q[3] << 24);
                    }
                }
                pre[i][k] = w;
            }
        }
    };
    auto store_stage = [&](uint32_t *z) {
#pragma unroll
        for (int i = 0; i < QPT; i++) {
            if (threadIdx.x + i * XE_THREADS >= NQUAD) continue;
            // 4x4 byte transpose: o[b] = byte b of the four time steps
            const uint32_t lo01 = __byte_perm(pre[i][0], pre[i][1], 0x5140);
            const uint32_t hi01 = __byte_perm(pre[i][0], pre[i][1], 0x7362);
            const uint32_t lo23 = __byte_perm(pre[i][2], pre[i][3], 0x5140);
            const uint32_t hi23 = __byte_perm(pre[i][2], pre[i][3], 0x7362);
            uint32_t *d = z + qz[i];
            d[0] = __byte_perm(lo01, lo23, 0x5410);                    // first sample, re
            d[8 * XE_RSW] = __byte_perm(lo01, lo23, 0x7632);           //               im
            d[ZB] = __byte_perm(hi01, hi23, 0x5410);                   // second sample, re
            d[ZB + 8 * XE_RSW] = __byte_perm(hi01, hi23, 0x7632);      //                im
        }
    };

    load_stage(s0);
    store_stage(zbuf);
    if (s0 + 1 < s1) load_stage(s0 + 1);
    xe_cta_sync();

    for (int sg = s0; sg < s1; sg++) {
        {
            const uint32_t *z = zbuf + ((sg - s0) & 1) * ZW + chl * CSW;
            if constexpr (NT > 0) {
                int a[MMAX + 1][4];
#pragma unroll
                for (int m = 0; m <= MMAX; m++) {
                    const uint32_t *r = z + (m * 16 + g) * XE_RSW + tig;
                    a[m][0] = (int)r[0];
                    a[m][1] = (int)r[8 * XE_RSW];
                    a[m][2] = (int)r[4];
                    a[m][3] = (int)r[8 * XE_RSW + 4];
                }
                static_for<0, WPC>([&](auto q_) {
                    constexpr int Q = decltype(q_)::value;
                    if (q != Q) return;
                    static_for<0, MT>([&](auto mi_) {
                        constexpr int mi = decltype(mi_)::value;
                        if constexpr (owns<WPC, Q>(mi)) {
                            constexpr int tb = tile_base<MT, WPC, Q>(mi);
                            static_for<0, mi + 1>([&](auto nj_) {
                                constexpr int nj = decltype(nj_)::value;
                                mma_s8(acc[tb + nj][0], a[mi], a[nj][0], a[nj][2]);   // columns = re rows of tile nj
                                mma_s8(acc[tb + nj][1], a[mi], a[nj][1], a[nj][3]);   // columns = im rows
                            });
                        }
                    });
                });
            }
            // feed: next stage into the other Z buffer, the one after it into registers
            if (sg + 1 < s1) {
                store_stage(zbuf + ((sg + 1 - s0) & 1) * ZW);
                if (sg + 2 < s1) load_stage(sg + 2);
            }
        }
        const int grp = sg / nst;
        const bool group_done = (sg + 1 == s1) || ((sg + 1) % nst == 0);
        if (!group_done) {
            xe_cta_sync();
            continue;
        }
        const int f0 = grp * FC;

        // ---- epilogue: combine re/im products, scatter the lower triangle ----
        // thread (g, tig) of row tile mi / column tile nj holds, for h = 0,1:
        //   v1 = 8 mi + g (row input), v2 = 8 nj + 2 tig + h (column input)
        // 1 pol : s = v;            o = f nbl + tri(s1) + s2
        // 2 pol : s = v>>1, p = v&1; o = 4 (f nbl + tri(s1) + s2) + 2 p1 + p2,  s2 = 4 nj + tig, p2 = h
        // off-diagonal tiles (nj < mi) are always inside the triangle
        const int f = f0 + chl;
        if constexpr (NT > 0) {
            if (f < p.F) {
              static_for<0, WPC>([&](auto q_) {
                constexpr int Q = decltype(q_)::value;
                if (q != Q) return;
                static_for<0, MT>([&](auto mi_) {
                    constexpr int mi = decltype(mi_)::value;
                    if constexpr (owns<WPC, Q>(mi)) {
                        constexpr int tb = tile_base<MT, WPC, Q>(mi);
                        const int v1 = mi * 8 + g;
                        const int s1 = (NPOL == 1) ? v1 : (v1 >> 1);
                        const bool row_ok = s1 < p.A;
                        long ob;        // output index of column input v2 = 2 tig, nj = 0
                        if (NPOL == 1) ob = (long)f * nbl + s1 * (s1 + 1) / 2 + 2 * tig;
                        else ob = 4 * ((long)f * nbl + s1 * (s1 + 1) / 2 + tig) + 2 * (v1 & 1);
                        static_for<0, mi + 1>([&](auto nj_) {
                            constexpr int nj = decltype(nj_)::value;
                            const int (&C)[4] = acc[tb + nj][0];
                            const int (&D)[4] = acc[tb + nj][1];
#pragma unroll
                            for (int h = 0; h < 2; h++) {
                                const int re = C[h] + D[2 + h];
                                const int im = C[2 + h] - D[h];
                                bool ok = row_ok;
                                if (nj == mi) ok = ok && ((NPOL == 1) ? (2 * tig + h <= g) : (tig <= (g >> 1)));
                                const long o = ob + ((NPOL == 1) ? (8 * nj + h) : (16 * nj + h));
                                if (ok) {
                                    if (p.split) {
                                        atomicAdd(p.out_i32 + 2 * o, re);
                                        atomicAdd(p.out_i32 + 2 * o + 1, im);
                                    } else if (p.out_i32) {
                                        int2 v = make_int2(re, im);
                                        if (p.accumulate) {
                                            int2 old = reinterpret_cast<int2 *>(p.out_i32)[o];
                                            v.x += old.x;
                                            v.y += old.y;
                                        }
                                        reinterpret_cast<int2 *>(p.out_i32)[o] = v;
                                    }
                                    if (!p.split && p.out_f32) {
                                        float2 v = make_float2((float)re * p.scale, (float)im * p.scale);
                                        if (p.accumulate) {
                                            float2 old = p.out_f32[o];
                                            v.x += old.x;
                                            v.y += old.y;
                                        }
                                        p.out_f32[o] = v;
                                    }
                                }
                            }
                        });
                    }
                });
              });
            }
        }
        zero_acc();
        xe_cta_sync();
    }
}

template <int MT, int WPC, int NPOL>
__global__ void __launch_bounds__(XE_THREADS, 1) k_xengine_i8(XeParams p)
{
    extern __shared__ __align__(16) uint32_t xe_smem[];
    xe_body<MT, WPC, NPOL>(p, xe_smem, (threadIdx.x >> 5) % WPC);
}

__device__ __forceinline__ void fence_proxy_async_smem()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

} // namespace
#include "xengine_tc.cuh"
#include "xengine_tma.cuh"
namespace {

// packed 4-bit (hi nibble re, lo nibble im) -> int8 pairs; LUT of CharToComplex (:833)
__global__ void k_unpack4(const uint8_t *__restrict__ in, int8_t *__restrict__ out, long n)
{
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int b = in[i];
        int hi = b >> 4, lo = b & 15;
        hi = hi < 8 ? hi : (hi == 8 ? 0 : hi - 16);
        lo = lo < 8 ? lo : (lo == 8 ? 0 : lo - 16);
        reinterpret_cast<char2 *>(out)[i] = make_char2((char)hi, (char)lo);
    }
}

// ---- complex-float input (the reference's default DTYPE_COMPLEX) ---------------------------------------------
// Per channel V = X X^H with X[(station, pol)][t] complex float: a CTA stages a tile of C32_TT time steps x NV
// rows x C32_CH channels in shared memory (transposed so that a thread's 4 rows / 4 columns are one 32 B read each)
// and every thread owns one 4 x 4 block of the LOWER triangle of one channel's matrix in registers: 64 FFMA per
// 8 shared-memory loads and time step.  The sum over t runs in order inside a thread, like the reference
// work-item's loop (lib/clXEngine_impl.cc:739-810, cxmac :729-736).  Few channel groups: the integration is also
// split into time slices whose partial matrices are summed by a second kernel in a fixed order (deterministic).
constexpr int C32_TT = 8;       // time steps per tile

struct XeC32 {
    const float2 *in;           // [t][station][Fstride][npol]
    float2 *out;                // [F][nbl][npol^2] (ts == 1) or partial matrices [ts][F*nbl*npol^2]
    int A, npol, NV, NB, nblk;  // NV = A*npol rows, NB = ceil(NV/4) blocks per side, nblk = NB(NB+1)/2
    int F, Fstride, f_off, T, ts, accumulate;
    int ch;                     // channels per CTA: ch * nblk <= 256 threads
    long nout;
};

constexpr int C32_SLOTS = 4;    // 16 B staging loads per thread and tile (vector path)

template <bool VEC>
__global__ void __launch_bounds__(256, 2) k_xengine_c32_tiled(XeC32 p)
{
    extern __shared__ __align__(16) float2 c32_smem[];          // 2 x [C32_TT][ch][NVP]
    const int NVP = p.NB * 4 + 2;                               // rows padded to whole blocks, + 16 B: the channels of a tile
                                                                // start on different banks (rows of 256 B collided)
    const int CH = p.ch;
    const int tid = threadIdx.x;
    const int grp = blockIdx.x, slice = blockIdx.y;
    const int f0 = grp * CH;
    const int c = tid / p.nblk, b = tid - c * p.nblk;           // channel of the group, block of the triangle
    const bool worker = c < CH;
    int bi = 0, bj = 0;
    if (worker) {
        bi = (int)((sqrtf(8.0f * b + 1.0f) - 1.0f) * 0.5f);
        while ((bi + 1) * (bi + 2) / 2 <= b) bi++;
        while (bi * (bi + 1) / 2 > b) bi--;
        bj = b - bi * (bi + 1) / 2;
    }
    const int tlen = (p.T + p.ts - 1) / p.ts;
    const int t0 = slice * tlen, t1 = min(p.T, t0 + tlen);
    float2 acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = make_float2(0.f, 0.f);

    const long frame = (long)p.A * p.Fstride * p.npol;          // float2 per time step
    const int per_t = p.NV * CH;                                // values per time step of the tile
    const int tile = C32_TT * CH * NVP;                         // float2 per staging buffer
    // padded rows read as zero in both buffers, once
    for (int e = tid; e < 2 * tile; e += blockDim.x) c32_smem[e] = make_float2(0.f, 0.f);

    // ---- vector staging: a thread's 16 B slots are the same (t, station, pair) in every tile ----
    // a (t, station) run holds CH * npol consecutive values = q2 float4; one float4 = two channels of one
    // polarisation (npol 1) or the X, Y values of one channel (npol 2)
    const int q2 = CH * p.npol / 2;
    int sl_t[C32_SLOTS], sl_s0[C32_SLOTS], sl_s1[C32_SLOTS];
    long sl_g[C32_SLOTS];
    bool sl_ok0[C32_SLOTS], sl_ok1[C32_SLOTS];
    if (VEC) {
#pragma unroll
        for (int k = 0; k < C32_SLOTS; k++) {
            const int e = tid + k * blockDim.x;
            const int t = e / (p.A * q2), r = e - t * (p.A * q2);
            const int s = r / q2, w = r - s * q2;
            const int ch0 = (p.npol == 1) ? 2 * w : w, ch1 = (p.npol == 1) ? 2 * w + 1 : w;
            const int v0 = (p.npol == 1) ? s : 2 * s, v1 = (p.npol == 1) ? s : 2 * s + 1;
            sl_t[k] = t;
            sl_g[k] = (long)t * frame + ((long)s * p.Fstride + p.f_off + f0) * p.npol + 2 * w;
            sl_s0[k] = (t * CH + ch0) * NVP + v0;
            sl_s1[k] = (t * CH + ch1) * NVP + v1;
            sl_ok0[k] = t < C32_TT && f0 + ch0 < p.F;
            sl_ok1[k] = t < C32_TT && f0 + ch1 < p.F;
        }
    }
    float4 pre[C32_SLOTS];
    auto load_tile = [&](int tb) {
        if (VEC) {
            const int nt = min(C32_TT, t1 - tb);
            const float2 *base = p.in + (long)tb * frame;
#pragma unroll
            for (int k = 0; k < C32_SLOTS; k++) {
                pre[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (sl_ok0[k] && sl_t[k] < nt) {
                    if (sl_ok1[k]) pre[k] = __ldg(reinterpret_cast<const float4 *>(base + sl_g[k]));
                    else {
                        const float2 v = __ldg(base + sl_g[k]);
                        pre[k] = make_float4(v.x, v.y, 0.f, 0.f);
                    }
                }
            }
        }
    };
    auto store_tile = [&](int tb, float2 *buf) {
        const int nt = min(C32_TT, t1 - tb);
        if (VEC) {
#pragma unroll
            for (int k = 0; k < C32_SLOTS; k++)
                if (sl_t[k] < nt && sl_ok0[k]) {
                    buf[sl_s0[k]] = make_float2(pre[k].x, pre[k].y);
                    buf[sl_s1[k]] = make_float2(pre[k].z, pre[k].w);      // (a channel past F: zeros into its own row)
                }
        } else {
            // scalar path (odd run lengths / unaligned rows): global [t][station][chan][pol] -> [t][chan][v]
            for (int e = tid; e < nt * per_t; e += blockDim.x) {
                const int t = e / per_t, r = e - t * per_t;
                const int s = r / (CH * p.npol), q = r - s * (CH * p.npol);
                const int ch = q / p.npol, pol = q - ch * p.npol;
                float2 v = make_float2(0.f, 0.f);
                if (f0 + ch < p.F) v = __ldg(p.in + (long)(tb + t) * frame + ((long)s * p.Fstride + p.f_off + f0 + ch) * p.npol + pol);
                buf[(t * CH + ch) * NVP + s * p.npol + pol] = v;
            }
        }
    };

    __syncthreads();
    load_tile(t0);
    store_tile(t0, c32_smem);
    int cur = 0;
    for (int tb = t0; tb < t1; tb += C32_TT) {
        const int nt = min(C32_TT, t1 - tb);
        const bool more = tb + C32_TT < t1;
        if (more) load_tile(tb + C32_TT);                       // the next tile's loads fly during this tile's FMAs
        __syncthreads();                                        // tile `cur` is complete, tile `cur^1` is free
        if (worker) {
            const float2 *buf = c32_smem + cur * tile;
            const float4 *row = reinterpret_cast<const float4 *>(buf + c * NVP + bi * 4);
            const float4 *col = reinterpret_cast<const float4 *>(buf + c * NVP + bj * 4);
            const int tstride = CH * NVP / 2;                   // float4 per time step
#pragma unroll 8
            for (int t = 0; t < nt; t++) {
                const float4 r01 = row[t * tstride], r23 = row[t * tstride + 1];
                const float4 c01 = col[t * tstride], c23 = col[t * tstride + 1];
                const float2 a[4] = {make_float2(r01.x, r01.y), make_float2(r01.z, r01.w), make_float2(r23.x, r23.y), make_float2(r23.z, r23.w)};
                const float2 bb[4] = {make_float2(c01.x, c01.y), make_float2(c01.z, c01.w), make_float2(c23.x, c23.y), make_float2(c23.z, c23.w)};
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        // accum += a * conj(b) (cxmac, lib/clXEngine_impl.cc:729-736), four chained FMAs
                        acc[i][j].x = fmaf(a[i].x, bb[j].x, acc[i][j].x);
                        acc[i][j].x = fmaf(a[i].y, bb[j].y, acc[i][j].x);
                        acc[i][j].y = fmaf(a[i].y, bb[j].x, acc[i][j].y);
                        acc[i][j].y = fmaf(-a[i].x, bb[j].y, acc[i][j].y);
                    }
            }
        }
        if (more) store_tile(tb + C32_TT, c32_smem + (cur ^ 1) * tile);
        cur ^= 1;
    }
    if (!worker || f0 + c >= p.F) return;
    const int nbl = p.A * (p.A + 1) / 2, pp = p.npol * p.npol;
    float2 *out = p.out + (long)slice * p.nout + (long)(f0 + c) * nbl * pp;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int v1 = bi * 4 + i, v2 = bj * 4 + j;
            if (v1 >= p.NV || v2 >= p.NV) continue;
            const int s1 = v1 / p.npol, p1 = v1 - s1 * p.npol, s2 = v2 / p.npol, p2 = v2 - s2 * p.npol;
            if (s2 > s1) continue;
            const long o = (long)(s1 * (s1 + 1) / 2 + s2) * pp + p1 * p.npol + p2;
            float2 v = acc[i][j];
            if (p.ts == 1 && p.accumulate) {
                v.x += out[o].x;
                v.y += out[o].y;
            }
            out[o] = v;
        }
}

// partial matrices of the time slices, summed in slice order
__global__ void k_c32_reduce(const float2 *__restrict__ part, float2 *__restrict__ out, long n, int ts, int accumulate)
{
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float2 v = part[i];
        for (int s = 1; s < ts; s++) {
            v.x += part[s * n + i].x;
            v.y += part[s * n + i].y;
        }
        if (accumulate) {
            v.x += out[i].x;
            v.y += out[i].y;
        }
        out[i] = v;
    }
}

// Streaming ingest from page-locked ports: the SMs read the ports over PCIe themselves (zero-copy loads, 16 B per lane,
// thousands in flight) and write the [t][station][chan] integration buffer -- one launch per push instead of one
// 2-D DMA per station (cudaMemcpy2DAsync of 1-2 KiB rows reaches ~27 GB/s of the link's 55).
struct XeIngest {
    const char *port[64];
    char *dst;                  // first time step of this push in the device integration buffer
    long ntime;
    int nports, row16;          // 16 B units per (t, station) row of this handle's slab
    long src_pitch, src_off;    // bytes between two time steps of a port / offset of the slab inside a port item
    long frame;                 // bytes per time step in the device buffer
};
__global__ void __launch_bounds__(256) k_xe_ingest(XeIngest g)
{
    const long per_t = (long)g.nports * g.row16;              // 16 B units per time step
    const long total = per_t * g.ntime;
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const long t = i / per_t;
        const int r = (int)(i - t * per_t);
        const int s = r / g.row16, c = r - s * g.row16;
        const int4 v = __ldcs(reinterpret_cast<const int4 *>(g.port[s] + t * g.src_pitch + g.src_off) + c);
        reinterpret_cast<int4 *>(g.dst + t * g.frame + (long)s * g.row16 * 16)[c] = v;
    }
}

// completion of the fused gather, stream-ordered behind the correlation kernel (whose peer stores have all been
// performed when it completes): lane r releases THIS rank's flag on rank r -- or lane 0 on every rank at once through
// the multicast address -- then acquires rank r's flag in the local array until it carries `epoch` (or a later one).
// What follows on the stream reads a complete matrix; no host barrier, no fence inside the hot kernel.
__global__ void k_gather_signal_wait(XeGatherSync g)
{
    // launched with programmatic stream serialisation: it is resident before the correlation kernel ends (its launch
    // latency is off the critical path) and lets the NEXT correlation kernel run its prologue meanwhile
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");            // the correlation kernel and its peer stores are complete
    const int r = threadIdx.x;
    if (g.signal) {
        if (g.flag_mc != nullptr) {
            if (r == 0) asm volatile("multimem.st.release.sys.global.u32 [%0], %1;" ::"l"(g.flag_mc), "r"(g.epoch) : "memory");
        } else if (r < g.nranks && g.flag[r] != nullptr) {
            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(g.flag[r]), "r"(g.epoch) : "memory");
        }
    }
    if (g.wait && r < g.nranks) {
        const unsigned *f = g.local + (size_t)r * CLB200_XENGINE_FLAG_STRIDE;
        unsigned v;
        unsigned long long t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > 10000000000ull) __trap();       // a rank that never signals: fail (CUDA error) instead of hanging
        } while ((int)(v - g.epoch) < 0);
    }
}

__global__ void k_i32_to_f32(const int2 *__restrict__ in, float2 *__restrict__ out, long n,
                             float scale, int accumulate)
{
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float2 v = make_float2((float)in[i].x * scale, (float)in[i].y * scale);
        if (accumulate) {
            v.x += out[i].x;
            v.y += out[i].y;
        }
        out[i] = v;
    }
}

// ---------------------------------------------------------------------- host --
typedef void (*xe_kernel_t)(XeParams);
struct XeVariant {
    int mt, wpc, fc, smem_bytes;
    xe_kernel_t kernel[2];          // [npol - 1]
};
template <int MT, int WPC>
XeVariant make_xe()
{
    constexpr int FC = XE_WARPS / WPC;
    constexpr int CSW = MT * 16 * XE_RSW + 2;
    return XeVariant{MT, WPC, FC, 2 * FC * CSW * 4, {&k_xengine_i8<MT, WPC, 1>, &k_xengine_i8<MT, WPC, 2>}};
}
const XeVariant *pick_xe(int nv)
{
    static const XeVariant tab[] = {
        make_xe<1, 1>(), make_xe<2, 1>(), make_xe<3, 1>(), make_xe<4, 1>(),
        make_xe<5, 2>(), make_xe<6, 2>(), make_xe<7, 2>(), make_xe<8, 4>(),
    };
    int mt = (nv + 7) / 8;
    if (mt < 1 || mt > 8) return nullptr;
    return &tab[mt - 1];
}

struct XEngine : clb200_block {
    int data_type = 0, npol = 1, A = 0, F = 0, T = 0;
    int Ftotal = 0, f_first = 0;       // channel shard within the caller's buffer
    const XeVariant *var = nullptr;
    bool use_tc = false;
    bool use_tma = false;              // TMA-fed variant of the tcgen05 kernel (needs 16 B aligned rows)
    bool use_tma_pk = false;           // ... and its packed 4-bit variant
    int l2promo = 0;
    void *gather[8] = {};              // full matrices of every rank (clb200_xengine_set_gather)
    int ngather = 0;
    // completion flags of the fused gather (clb200_xengine_set_gather_sync)
    int gather_rank = -1;
    unsigned *gather_flags[8] = {};    // flag array of every rank (peer-mapped): word [src * FLAG_STRIDE] = epoch of src's slab
    void *gather_mc = nullptr;         // NVSwitch multicast address of the matrix (optional)
    unsigned *gather_flags_mc = nullptr;
    unsigned gather_epoch = 0;
    Buf d_gather_counter;
    int fc_override = 0;               // CLB200_XE_FC: channels per CTA of the TMA kernel (8 | 16)
    bool pdl = true;                   // programmatic dependent launch (CLB200_XE_PDL=0 turns it off)
    Buf d_in[2], d_unpacked, d_acc, d_out;
    Buf pin_in[2], pin_out;
    cudaStream_t s_copy = nullptr, s_comp = nullptr;
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr}, ev_done = nullptr;
    // ---- streaming ingest (clb200_xengine_stream_begin / push_timesteps / poll_result) ----
    static constexpr int MAXRES = 8;
    struct Stream {
        bool on = false;
        int pipeline = 1, nres = 4;
        Buf pin[2], dev[2];                    // the integration being filled / the one being correlated
        Buf d_res[MAXRES], pin_res[MAXRES];    // result ring: device matrix + its pinned host copy
        cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
        cudaEvent_t ev_h2d[2] = {}, ev_kern[2] = {};          // uploads of buffer b landed / kernel on buffer b done
        cudaEvent_t ev_kres[MAXRES] = {}, ev_res[MAXRES] = {}; // matrix r computed / copied to the host
        bool pin_dirty[2] = {false, false};    // pinned staging b still has uploads in flight
        bool dma_ingest = false;               // CLB200_XE_INGEST_DMA=1: one 2-D DMA per station instead of the ingest kernel
        bool ports_stable = false;             // page-locked ports stay untouched until their integration's result is polled
        long tracker = 0;                      // time steps of the current integration ingested so far
        long n_integ = 0;                      // integrations handed to the GPU
        long res_head = 0, res_tail = 0;       // results launched / delivered
        int pipe_count = 0;
        uint64_t n_push = 0, n_blocked = 0;    // push calls / calls that had to wait for the GPU
    } st;
    long nbl() const { return (long)A * (A + 1) / 2; }
    long out_items() const { return (long)F * nbl() * npol * npol; }
    size_t sample_bytes() const
    {
        return data_type == CLB200_DTYPE_COMPLEX ? 8 : (data_type == CLB200_DTYPE_BYTE ? 2 : 1);
    }
    ~XEngine() override
    {
        DeviceGuard g(device);
        if (s_copy) cudaStreamSynchronize(s_copy);
        if (s_comp) cudaStreamSynchronize(s_comp);
        for (int i = 0; i < 2; i++) {
            d_in[i].release();
            pin_in[i].release();
            if (ev_in[i]) cudaEventDestroy(ev_in[i]);
            if (ev_free[i]) cudaEventDestroy(ev_free[i]);
        }
        d_unpacked.release();
        d_acc.release();
        d_out.release();
        pin_out.release();
        d_gather_counter.release();
        if (ev_done) cudaEventDestroy(ev_done);
        stream_free();
        if (s_copy) cudaStreamDestroy(s_copy);
        if (s_comp) cudaStreamDestroy(s_comp);
    }
    void stream_free()
    {
        if (st.s_h2d) cudaStreamSynchronize(st.s_h2d);
        if (s_comp) cudaStreamSynchronize(s_comp);
        if (st.s_d2h) cudaStreamSynchronize(st.s_d2h);
        for (int b = 0; b < 2; b++) {
            st.pin[b].release();
            st.dev[b].release();
            if (st.ev_h2d[b]) cudaEventDestroy(st.ev_h2d[b]);
            if (st.ev_kern[b]) cudaEventDestroy(st.ev_kern[b]);
        }
        for (int r = 0; r < MAXRES; r++) {
            st.d_res[r].release();
            st.pin_res[r].release();
            if (st.ev_kres[r]) cudaEventDestroy(st.ev_kres[r]);
            if (st.ev_res[r]) cudaEventDestroy(st.ev_res[r]);
        }
        if (st.s_h2d) cudaStreamDestroy(st.s_h2d);
        if (st.s_d2h) cudaStreamDestroy(st.s_d2h);
        st = Stream();
    }
};

// enqueue the correlation of `T` time steps held at d_in (layout [t][A][Fstride][npol]);
// exactly one of out_i32 / out_f32 may be null
int xe_launch(XEngine *x, const void *d_in, int T, int Fstride, int f_off, int32_t *out_i32,
              float2 *out_f32, int accumulate, cudaStream_t st, bool gather = false, int nbatch = 1)
{
    const int sms = device_sm_count(x->device);
    if (x->data_type == CLB200_DTYPE_COMPLEX) {
        CLB_CHECK(out_f32 != nullptr, CLB200_EINVAL, "clXEngine: complex input has no integer output");
        XeC32 q;
        q.in = (const float2 *)d_in;
        q.A = x->A;
        q.npol = x->npol;
        q.NV = x->A * x->npol;
        q.NB = (q.NV + 3) / 4;
        q.nblk = q.NB * (q.NB + 1) / 2;
        q.F = x->F;
        q.Fstride = Fstride;
        q.f_off = f_off;
        q.T = T;
        q.accumulate = accumulate;
        q.nout = x->out_items();
        CLB_CHECK(q.nblk <= 256, CLB200_EINVAL, "clXEngine: too many inputs for the complex kernel");
        // channels per CTA: one thread per (channel, 4 x 4 block), at most 256 threads and 48 KiB of staged samples
        int ch = std::min(std::min(256 / q.nblk, 384 / (q.NB * 4 + 2)), 32);
        if (ch >= 4) ch &= ~3;
        ch = std::max(1, std::min(ch, x->F));
        q.ch = ch;
        const int groups = (x->F + ch - 1) / ch;
        const int threads = (ch * q.nblk + 31) / 32 * 32;
        // enough CTAs for ~4 per SM: split the integration into time slices when there are few channel groups
        static const int per_sm = [] {
            const char *e = getenv("CLB200_XE_C32_CTAS");       // tuning
            return e && atoi(e) > 0 ? atoi(e) : 4;
        }();
        int ts = std::max(1, std::min((per_sm * sms + groups - 1) / groups, (T + 4 * C32_TT - 1) / (4 * C32_TT)));
        ts = std::min(ts, 64);
        q.ts = ts;
        if (ts > 1) {
            CLB_TRY(x->d_acc.reserve((size_t)ts * q.nout * 8));
            q.out = (float2 *)x->d_acc.p;
        } else {
            q.out = out_f32;
        }
        const size_t smem = 2 * (size_t)C32_TT * ch * (q.NB * 4 + 2) * sizeof(float2);   // two staging buffers
        // 16 B staging loads: whole float4 per (t, station) run, 16 B aligned rows, few enough slots per thread
        const bool vec = (ch * x->npol) % 2 == 0 && ((long)Fstride * x->npol) % 2 == 0 && ((long)f_off * x->npol) % 2 == 0 &&
                         ((uintptr_t)d_in % 16) == 0 && (ch % 2 == 0 || x->npol == 2) &&
                         (long)C32_TT * x->A * (ch * x->npol / 2) <= (long)C32_SLOTS * threads;
        if (vec) k_xengine_c32_tiled<true><<<dim3(groups, ts), threads, smem, st>>>(q);
        else k_xengine_c32_tiled<false><<<dim3(groups, ts), threads, smem, st>>>(q);
        CLB_CUDA(cudaGetLastError());
        x->n_launch++;
        if (ts > 1) {
            k_c32_reduce<<<grid_for((q.nout + 255) / 256, sms, 8), 256, 0, st>>>((const float2 *)x->d_acc.p, out_f32, q.nout, ts,
                                                                               accumulate);
            x->n_launch++;
        }
        CLB_CUDA(cudaGetLastError());
        x->n_launch++;
        return CLB200_OK;
    }
    const int8_t *src = (const int8_t *)d_in;
    float scale = 1.0f / (127.0f * 127.0f);
    // packed 4 bit: the TMA kernel expands the nibbles in its transpose stage when the packed rows are 16 B aligned
    // (and one sample lane run is >= 16 B: 16 channels per CTA); otherwise a separate pass unpacks into int8 pairs
    bool pk = false;
    if (x->data_type == CLB200_DTYPE_PACKEDXY) {
        pk = x->use_tc && x->use_tma && x->use_tma_pk && ((uintptr_t)d_in % 16 == 0) &&
             (((long)Fstride * x->npol) % 16 == 0) && ((uintptr_t)out_i32 % 8 == 0) && ((uintptr_t)out_f32 % 8 == 0) &&
             (((long)f_off * x->npol) % 16 == 0) && nbatch == 1;
        scale = 1.0f / 49.0f;
    }
    if (x->data_type == CLB200_DTYPE_PACKEDXY && !pk) {
        long n = (long)T * x->A * Fstride * x->npol;
        CLB_TRY(x->d_unpacked.reserve((size_t)n * 2));
        k_unpack4<<<grid_for((n + 255) / 256, sms, 8), 256, 0, st>>>((const uint8_t *)d_in,
                                                                     (int8_t *)x->d_unpacked.p, n);
        CLB_CUDA(cudaGetLastError());
        x->n_launch++;
        src = (const int8_t *)x->d_unpacked.p;
    }
    const XeVariant *v = x->var;
    const bool tc = x->use_tc;                         // tcgen05/TMEM kernel (<= 32 rows of inputs)
    const long nout = x->out_items();
    long rowb = (long)Fstride * x->npol * (pk ? 1 : 2);
    const bool tma_ok = tc && x->use_tma && ((uintptr_t)src % 16 == 0) && (rowb % 16 == 0) &&
                        ((uintptr_t)out_i32 % 8 == 0) && ((uintptr_t)out_f32 % 8 == 0);
    // TMA kernel: 8 channels per CTA while 16 would leave SMs without a channel group (then no CTA
    // has to share a group with another one: no time slicing, no cross-CTA reduction)
    int fc = tc ? TC_FC : v->fc;
    int want_slices = -1;                                // -1: derive from the group count below
    if (tma_ok) {
        // (channels per CTA, time slices per group) = the pair that puts the most CTAs on the SMs in ONE wave with
        // clusters of at most 4 (measured, tools/xe_small.py: 8-CTA clusters cost 24.9 vs 12.2 us at 128 channels,
        // 46 vs 17 us at 512).  Ties: 16 channels (32 B runs stream at 5.9 instead of 3.7 TB/s) when a CTA streams
        // >= 384 KiB, else 8 channels with half the slices (F = 512: 15.9 vs 17.1 us).
        long best = -1;
        fc = 16;
        const long in_bytes = (long)T * x->A * x->F * x->npol * 2;
        for (int cand = 16; cand >= 8; cand -= 8) {
            const int ng = (x->F + cand - 1) / cand, nstc = (T + 512 / cand - 1) / (512 / cand);
            for (int sl = 1; sl <= 4; sl *= 2) {
                if (sl > nstc || (long)ng * sl > sms) continue;
                const long ctas = (long)ng * sl;
                bool take = ctas > best;
                if (ctas == best && cand == 8) take = in_bytes / ctas < 384 * 1024;
                if (take) {
                    best = ctas;
                    fc = cand;
                    want_slices = sl;
                }
            }
        }
        if (best < 0) want_slices = -1;                  // more groups than SMs either way: whole groups per CTA, 16 channels
        // two polarisations: 8 channels are already 32 B runs, and enough 8-channel groups to fill the SMs need
        // no time slicing, i.e. no exchange at all (measured 18.2 vs 20.7 us at 16 stations x 2 pols x 1024 channels)
        if (x->npol == 2 && (x->F + 7) / 8 >= sms / 2 && (x->F + 7) / 8 <= sms) {
            fc = 8;
            want_slices = 1;
        }
        if (nbatch > 1 || pk) fc = 16;                   // persistent over (integration, group): the widest rows; packed: >= 16 B runs
        if (pk && want_slices > 0) {                     // the slice count that goes with 16 channels
            const int ng = (x->F + 15) / 16, nstc = (T + 31) / 32;
            want_slices = 1;
            while (want_slices * 2 <= 4 && want_slices * 2 <= nstc && (long)ng * want_slices * 2 <= sms) want_slices *= 2;
        }
        if ((x->fc_override == 8 || x->fc_override == 16) && !pk) {
            fc = x->fc_override;
            want_slices = -1;
        }
    }
    const int kt = tma_ok ? 512 / fc : XE_TT;           // time steps per stage
    const int ngroups = (x->F + fc - 1) / fc;
    // fewer channel groups than ~2 waves of CTAs: split the integrations over time as well
    const int nst = (T + kt - 1) / kt;
    // few channel groups: one wave of (slice, group) CTAs; many: whole groups per CTA
    int tslices = 0;
    if (ngroups < sms && nst > 1) tslices = std::min(nst, std::max(1, sms / ngroups));
    if (want_slices > 0) tslices = want_slices;
    {
        const char *e = getenv("CLB200_XE_SLICES");       // tuning override: 0 = stream-K split
        if (e && ngroups < 2 * sms && nst > 1) tslices = std::min(nst, atoi(e));
    }
    if (gather && tslices == 0) tslices = 1;            // whole groups per CTA: no partial sums in memory
    // batches: persistent CTAs over (integration, channel group) pairs, whole integrations per pair -- no time slicing,
    // no exchange, and the pipeline runs across pair boundaries (CLB200_XE_BATCH_SLICED=1 keeps the sliced grid x K)
    static const bool batch_sliced = [] { const char *e = getenv("CLB200_XE_BATCH_SLICED"); return e && atoi(e); }();
    if (nbatch > 1 && tma_ok && !batch_sliced) tslices = 0;
    else if (nbatch > 1 && tslices == 0) tslices = 1;
    if (tma_ok && tslices > 1) {
        // the slices of a group are the CTAs of one cluster: 2 or 4 (8 only on request), each finalising fc/slices channels
        const int cap = getenv("CLB200_XE_SLICES") ? 8 : 4;
        int c = 2;
        while (c * 2 <= std::min(tslices, std::min(cap, fc))) c *= 2;
        tslices = c;
    }
    const bool split = (tslices > 1) || (tslices == 0 && nbatch == 1 && ngroups < 2 * sms && nst > 1);
    const int nslice = split ? 2 : 1;
    int grid = tslices > 0 ? ngroups * tslices
                           : (int)std::min<long>(sms, split ? (long)ngroups * nst : (long)ngroups * nbatch);
    if (nbatch > 1 && tslices == 0) {
        static const int bg = [] { const char *e = getenv("CLB200_XE_BATCH_GRID"); return e ? atoi(e) : 0; }();   // tuning
        if (bg > 0) grid = std::min(bg, ngroups * nbatch);
    }
    CUtensorMap tmap;
    const bool tma = tma_ok && tm_make_map(&tmap, src, rowb, x->A, T, x->npol, fc, x->l2promo, nbatch, pk);
    CLB_CHECK(tma || nbatch == 1, CLB200_ESTATE, "clXEngine: batched launches need the TMA kernel");
    CLB_CHECK(tma || !tma_ok, CLB200_ECUDA, "clXEngine: cuTensorMapEncodeTiled failed");
    // TMA kernel, time-sliced: the slices of a channel group form a thread-block cluster and meet
    // through distributed shared memory -- no memset, no atomics, no conversion pass
    const bool fixup = tma && tslices > 1;
    if (!fixup && nslice > 1) {
        // partial sums meet in an int32 buffer through atomics
        if (out_i32 == nullptr) {
            CLB_TRY(x->d_acc.reserve((size_t)nout * 8));
            CLB_CUDA(cudaMemsetAsync(x->d_acc.p, 0, (size_t)nout * 8, st));
        } else if (!accumulate) {
            CLB_CUDA(cudaMemsetAsync(out_i32, 0, (size_t)nout * 8, st));
        }
    }
    XeParams p;
    p.in = src;
    p.out_i32 = (!fixup && nslice > 1 && out_i32 == nullptr) ? (int32_t *)x->d_acc.p : out_i32;
    p.out_f32 = out_f32;
    p.split = split ? 1 : 0;
    p.nslice = tslices;
    p.nbatch = nbatch;
    {
        const char *e = getenv("CLB200_XE_L2ROWS");
        p.l2_rows = e ? atoi(e) : 0;      // measured slower on B200 (50.9 vs 44.7 us): off by default
    }
    p.A = x->A;
    p.npol = x->npol;
    p.F = x->F;
    p.Fstride = Fstride;
    p.f_off = f_off;
    p.T = T;
    p.accumulate = accumulate;
    p.ngather = 0;
    p.gather_off = 0;
    p.gather_mc = nullptr;
    p.gather_flag_mc = nullptr;
    p.gather_counter = nullptr;
    p.gather_fence_gpu = 0;
    {
        // CLB200_XE_PDL_NOWAIT=1 (measurement only): consecutive launches skip griddepcontrol.wait, so their CTAs
        // interleave as SMs free up (21.6 -> 18.7 us at 1024 channels, 12.2 -> 6.8 us at 128).  NOT the default: the
        // launch would then also run ahead of whatever kernel produced its input on the same stream (a device-resident
        // F-engine -> X-engine chain), which the library cannot see; batches of independent integrations have
        // clb200_xengine_launch_device_batch instead.
        static const bool nowait = [] { const char *e = getenv("CLB200_XE_PDL_NOWAIT"); return e && atoi(e); }();
        p.pdl_nowait = (nowait && x->pdl && !accumulate && !gather) ? 1 : 0;
    }
    p.gather_epoch = 0;
    for (int r = 0; r < 8; r++) p.gather_flag[r] = nullptr;
    if (gather && x->gather_rank >= 0) {
        p.gather_mc = (float2 *)x->gather_mc;
        p.gather_epoch = ++x->gather_epoch;
        // CLB200_XE_GATHER_INKERNEL=1: the kernel's last CTA releases the flags itself (every CTA then waits for the
        // acknowledgement of its peer stores before it leaves its SM: measured 27.8 vs ... us per launch at 2 GPUs);
        // default: the flags are released by the one-warp kernel that clb200_xengine_gather_wait enqueues behind it
        static const int inkernel = [] { const char *e = getenv("CLB200_XE_GATHER_INKERNEL"); return e ? atoi(e) : 0; }();
        p.gather_fence_gpu = inkernel == 2;      // 2: CTAs release at GPU scope, only the last CTA fences system-wide
        if (inkernel) {
            p.gather_counter = (unsigned *)x->d_gather_counter.p;
            if (x->gather_flags_mc) p.gather_flag_mc = x->gather_flags_mc + (size_t)x->gather_rank * CLB200_XENGINE_FLAG_STRIDE;
            for (int r = 0; r < x->ngather; r++)
                p.gather_flag[r] = x->gather_flags[r] ? x->gather_flags[r] + (size_t)x->gather_rank * CLB200_XENGINE_FLAG_STRIDE : nullptr;
        }
    }
    if (gather) {
        CLB_CHECK(tma, CLB200_ESTATE, "clXEngine: the peer-memory gather needs the TMA kernel (16 B aligned rows, <= 32 inputs x pols)");
        p.ngather = x->ngather;
        for (int r = 0; r < x->ngather; r++) p.gather[r] = (float2 *)x->gather[r];
        p.gather_off = (long)x->f_first * x->nbl() * x->npol * x->npol;
    }
    p.scale = scale;
    p.aligned = ((uintptr_t)src % 4 == 0) && (rowb % 4 == 0) && (((long)f_off * x->npol * 2) % 4 == 0);
    if (tma) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(tslices > 0 ? grid * nbatch : grid);
        cfg.blockDim = dim3(TM_THREADS);
        cfg.dynamicSmemBytes = TM_SMEM;
        cfg.stream = st;
        cudaLaunchAttribute at[2];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = fixup ? tslices : 1;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at;
        cfg.numAttrs = x->pdl ? 2 : 1;
        CLB_CUDA(cudaLaunchKernelEx(&cfg, tm_kernel(x->npol, fc, pk), p, tmap));
    } else if (tc) {
        if (x->npol == 1) k_xengine_tc<1><<<grid, TC_THREADS, TC_SMEM, st>>>(p);
        else k_xengine_tc<2><<<grid, TC_THREADS, TC_SMEM, st>>>(p);
    } else {
        v->kernel[x->npol - 1]<<<grid, XE_THREADS, v->smem_bytes, st>>>(p);
    }
    CLB_CUDA(cudaGetLastError());
    x->n_launch++;
    if (!fixup && nslice > 1 && out_f32 != nullptr) {
        k_i32_to_f32<<<grid_for((nout + 255) / 256, sms, 8), 256, 0, st>>>(
            (const int2 *)p.out_i32, out_f32, nout, scale, accumulate);
        CLB_CUDA(cudaGetLastError());
        x->n_launch++;
    }
    return CLB200_OK;
}

int xe_init_streams(XEngine *x)
{
    if (x->s_copy) return CLB200_OK;
    CLB_CUDA(cudaStreamCreateWithFlags(&x->s_copy, cudaStreamNonBlocking));
    CLB_CUDA(cudaStreamCreateWithFlags(&x->s_comp, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) {
        CLB_CUDA(cudaEventCreateWithFlags(&x->ev_in[i], cudaEventDisableTiming));
        CLB_CUDA(cudaEventCreateWithFlags(&x->ev_free[i], cudaEventDisableTiming));
        x->pin_in[i].host = true;
    }
    x->pin_out.host = true;
    CLB_CUDA(cudaEventCreateWithFlags(&x->ev_done, cudaEventDisableTiming));
    return CLB200_OK;
}

// host-buffer integration: time-chunked H2D on one stream overlapped with the
// correlation of the previous chunk on another; int32 partial sums stay on the device
int xe_work(XEngine *x, const void *in, void *out, bool want_i32, int accumulate)
{
    CLB_TRY(xe_init_streams(x));
    const int Ftot = x->Ftotal > 0 ? x->Ftotal : x->F;
    const size_t sb = x->sample_bytes();
    const size_t width = (size_t)x->F * x->npol * sb;             // bytes of this shard per (t, station)
    const size_t pitch = (size_t)Ftot * x->npol * sb;
    const long rows_per_t = x->A;
    const bool is_int = x->data_type != CLB200_DTYPE_COMPLEX;
    CLB_CHECK(!(want_i32 && !is_int), CLB200_EINVAL, "clXEngine: complex input has no integer output");
    const long nout = x->out_items();
    const bool pinned = is_pinned(in);
    // chunk: about 8 MiB of input, a multiple of the 32-step stage
    long tch = std::max<long>(XE_TT, ((long)(8 << 20) / std::max<size_t>(1, width * rows_per_t)) / XE_TT * XE_TT);
    tch = std::min<long>(tch, x->T);
    if (is_int) CLB_TRY(x->d_acc.reserve((size_t)nout * 8));
    CLB_TRY(x->d_out.reserve((size_t)nout * 8));
    int c = 0;
    for (long t0 = 0; t0 < x->T; t0 += tch, c++) {
        const long nt = std::min<long>(tch, x->T - t0);
        const int b = c & 1;
        CLB_TRY(x->d_in[b].reserve(width * rows_per_t * tch));
        if (c >= 2) CLB_CUDA(cudaStreamWaitEvent(x->s_copy, x->ev_free[b], 0));
        const char *src = (const char *)in + (size_t)t0 * rows_per_t * pitch + (size_t)x->f_first * x->npol * sb;
        if (!pinned) {
            // stage through pinned memory (only the shard's columns)
            if (c >= 2) CLB_CUDA(cudaEventSynchronize(x->ev_in[b]));
            CLB_TRY(x->pin_in[b].reserve(width * rows_per_t * tch));
            for (long r = 0; r < nt * rows_per_t; r++)
                memcpy((char *)x->pin_in[b].p + r * width, src + r * pitch, width);
            CLB_CUDA(cudaMemcpyAsync(x->d_in[b].p, x->pin_in[b].p, width * rows_per_t * nt,
                                     cudaMemcpyHostToDevice, x->s_copy));
        } else {
            CLB_CUDA(cudaMemcpy2DAsync(x->d_in[b].p, width, src, pitch, width, nt * rows_per_t,
                                       cudaMemcpyHostToDevice, x->s_copy));
        }
        x->n_h2d += width * rows_per_t * nt;
        CLB_CUDA(cudaEventRecord(x->ev_in[b], x->s_copy));
        CLB_CUDA(cudaStreamWaitEvent(x->s_comp, x->ev_in[b], 0));
        if (is_int)
            CLB_TRY(xe_launch(x, x->d_in[b].p, (int)nt, x->F, 0, (int32_t *)x->d_acc.p, nullptr, c > 0,
                              x->s_comp));
        else
            CLB_TRY(xe_launch(x, x->d_in[b].p, (int)nt, x->F, 0, nullptr, (float2 *)x->d_out.p,
                              (c > 0) || accumulate, x->s_comp));
        CLB_CUDA(cudaEventRecord(x->ev_free[b], x->s_comp));
    }
    const void *res = x->d_out.p;
    if (is_int) {
        if (want_i32) {
            res = x->d_acc.p;
        } else {
            float scale = x->data_type == CLB200_DTYPE_PACKEDXY ? 1.0f / 49.0f : 1.0f / (127.0f * 127.0f);
            k_i32_to_f32<<<grid_for((nout + 255) / 256, device_sm_count(x->device), 8), 256, 0, x->s_comp>>>(
                (const int2 *)x->d_acc.p, (float2 *)x->d_out.p, nout, scale, accumulate);
            CLB_CUDA(cudaGetLastError());
            x->n_launch++;
        }
    }
    if (is_pinned(out)) {
        CLB_CUDA(cudaMemcpyAsync(out, res, (size_t)nout * 8, cudaMemcpyDeviceToHost, x->s_comp));
        CLB_CUDA(cudaStreamSynchronize(x->s_comp));
    } else {
        CLB_TRY(x->pin_out.reserve((size_t)nout * 8));
        CLB_CUDA(cudaMemcpyAsync(x->pin_out.p, res, (size_t)nout * 8, cudaMemcpyDeviceToHost, x->s_comp));
        CLB_CUDA(cudaStreamSynchronize(x->s_comp));
        memcpy(out, x->pin_out.p, (size_t)nout * 8);
    }
    x->n_d2h += (size_t)nout * 8;
    return CLB200_OK;
}

// ---- streaming ingest --------------------------------------------------------------------------
// The reference marshals every general_work() call into one of two pinned integration buffers and a worker
// thread uploads + correlates a FULL buffer while the scheduler thread fills the other one
// (work_processor lib/clXEngine_impl.cc:918-1142, runThread :1234-1299).  Here every push is DMA'd at
// once on a copy stream, so that when the last time step of an integration arrives only that last piece,
// the kernel and the result read-back remain; nothing in push() waits for the GPU unless the GPU is a
// whole integration behind (back-pressure), and results are picked up later with poll_result().
int xe_stream_begin(XEngine *x, int pipeline, int nres)
{
    XEngine::Stream &st = x->st;
    CLB_CHECK(!st.on, CLB200_ESTATE, "clXEngine: stream already begun");
    CLB_CHECK(nres >= 2 && nres <= XEngine::MAXRES, CLB200_EINVAL, "clXEngine: result slots must be 2..%d", XEngine::MAXRES);
    CLB_TRY(xe_init_streams(x));
    st.pipeline = pipeline > 1 ? pipeline : 1;
    st.nres = nres;
    {
        const char *e = getenv("CLB200_XE_INGEST_DMA");
        st.dma_ingest = e && atoi(e);
    }
    const size_t in_bytes = (size_t)x->T * x->A * x->F * x->npol * x->sample_bytes();      // this handle's slab
    const size_t out_bytes = (size_t)x->out_items() * 8;
    CLB_CUDA(cudaStreamCreateWithFlags(&st.s_h2d, cudaStreamNonBlocking));
    CLB_CUDA(cudaStreamCreateWithFlags(&st.s_d2h, cudaStreamNonBlocking));
    for (int b = 0; b < 2; b++) {
        st.pin[b].host = true;
        CLB_TRY(st.pin[b].reserve(in_bytes));
        CLB_TRY(st.dev[b].reserve(in_bytes));
        CLB_CUDA(cudaEventCreateWithFlags(&st.ev_h2d[b], cudaEventDisableTiming));
        CLB_CUDA(cudaEventCreateWithFlags(&st.ev_kern[b], cudaEventDisableTiming));
    }
    for (int r = 0; r < nres; r++) {
        st.pin_res[r].host = true;
        CLB_TRY(st.pin_res[r].reserve(out_bytes));
        CLB_TRY(st.d_res[r].reserve(out_bytes));
        CLB_CUDA(cudaEventCreateWithFlags(&st.ev_kres[r], cudaEventDisableTiming));
        CLB_CUDA(cudaEventCreateWithFlags(&st.ev_res[r], cudaEventDisableTiming));
    }
    if (x->data_type != CLB200_DTYPE_COMPLEX) CLB_TRY(x->d_acc.reserve(out_bytes));
    st.on = true;
    return CLB200_OK;
}

// hand the integration in buffer b to the GPU: kernel after its uploads, result read-back after the kernel
int xe_stream_launch(XEngine *x, int b)
{
    XEngine::Stream &st = x->st;
    const int r = (int)(st.res_head % st.nres);
    const bool first_of_group = st.pipe_count == 0;
    CLB_CUDA(cudaEventRecord(st.ev_h2d[b], st.s_h2d));
    CLB_CUDA(cudaStreamWaitEvent(x->s_comp, st.ev_h2d[b], 0));
    if (first_of_group && st.res_head >= st.nres)                       // matrix r is being rewritten: its last read-back is over
        CLB_CUDA(cudaStreamWaitEvent(x->s_comp, st.ev_res[r], 0));
    CLB_TRY(xe_launch(x, st.dev[b].p, x->T, x->F, 0, nullptr, (float2 *)st.d_res[r].p, first_of_group ? 0 : 1, x->s_comp));
    CLB_CUDA(cudaEventRecord(st.ev_kern[b], x->s_comp));
    st.n_integ++;
    if (++st.pipe_count >= st.pipeline) {                               // pipeline_integration (:785-808): emit every Nth
        st.pipe_count = 0;
        CLB_CUDA(cudaEventRecord(st.ev_kres[r], x->s_comp));
        CLB_CUDA(cudaStreamWaitEvent(st.s_d2h, st.ev_kres[r], 0));
        CLB_CUDA(cudaMemcpyAsync(st.pin_res[r].p, st.d_res[r].p, (size_t)x->out_items() * 8, cudaMemcpyDeviceToHost, st.s_d2h));
        CLB_CUDA(cudaEventRecord(st.ev_res[r], st.s_d2h));
        x->n_d2h += (size_t)x->out_items() * 8;
        st.res_head++;
    }
    return CLB200_OK;
}

// ports: the block's input streams for `ntime` time steps -- ports[s] holds ntime items of this station
// (row = all channels of the caller's stream); two polarisations of unpacked data arrive as separate ports
// X = ports[s], Y = ports[s + A] and are interleaved per channel (:1010-1057)
int xe_stream_push(XEngine *x, const void *const *ports, int nports, long ntime)
{
    XEngine::Stream &st = x->st;
    CLB_CHECK(st.on, CLB200_ESTATE, "clXEngine: stream_begin first");
    const bool planar_pol = x->npol == 2 && x->data_type != CLB200_DTYPE_PACKEDXY;
    const int need = planar_pol ? 2 * x->A : x->A;
    CLB_CHECK(nports == need, CLB200_EINVAL, "clXEngine: %d ports given, %d expected", nports, need);
    const int Ftot = x->Ftotal > 0 ? x->Ftotal : x->F;
    const size_t sb = x->sample_bytes();
    const size_t vec_in = (size_t)Ftot * sb * (planar_pol ? 1 : x->npol);     // bytes of one port item
    const size_t off_in = (size_t)x->f_first * sb * (planar_pol ? 1 : x->npol);
    const size_t row = (size_t)x->F * x->npol * sb;                            // one station, one time step (slab)
    const size_t frame = row * x->A;
    st.n_push++;
    long done = 0;
    while (done < ntime) {
        const int b = (int)((st.n_integ) & 1);
        const long n = std::min<long>(ntime - done, x->T - st.tracker);
        // zero-copy route: every port is page-locked (clb200_register_host_buffer) and no interleave is needed ->
        // the copy engine gathers the rows straight out of the ports; the call returns after the DMA has read them
        bool direct = !planar_pol && (size_t)n * frame >= ((size_t)1 << 20);
        for (int s = 0; direct && s < x->A; s++) direct = is_pinned(ports[s]);
        if (st.tracker == 0 && st.pipe_count == 0)
            // checked before anything of the integration is consumed, so that a refused push changes nothing
            CLB_CHECK(st.res_head - st.res_tail < st.nres, CLB200_ESTATE,
                      "clXEngine: %d results pending -- poll_result() before pushing more", st.nres);
        if (st.tracker == 0 && st.n_integ >= 2) {
            // buffer b was last used by integration n_integ-2: its kernel must be done before the copy stream
            // overwrites the device buffer (stream-ordered), and its uploads before the host overwrites the staging
            CLB_CUDA(cudaStreamWaitEvent(st.s_h2d, st.ev_kern[b], 0));
            if (st.pin_dirty[b]) {
                if (cudaEventQuery(st.ev_h2d[b]) != cudaSuccess) {
                    st.n_blocked++;
                    CLB_CUDA(cudaEventSynchronize(st.ev_h2d[b]));
                }
                st.pin_dirty[b] = false;
            }
        }
        char *dst_dev = (char *)st.dev[b].p + (size_t)st.tracker * frame;
        if (direct) {
            bool by_kernel = x->A <= 64 && row % 16 == 0 && vec_in % 16 == 0 && off_in % 16 == 0 && !st.dma_ingest;
            for (int s = 0; by_kernel && s < x->A; s++) by_kernel = ((uintptr_t)ports[s] % 16) == 0;
            if (by_kernel) {
                XeIngest g;
                for (int s = 0; s < x->A; s++) g.port[s] = (const char *)ports[s] + (size_t)done * vec_in;
                g.dst = dst_dev;
                g.ntime = n;
                g.nports = x->A;
                g.row16 = (int)(row / 16);
                g.src_pitch = (long)vec_in;
                g.src_off = (long)off_in;
                g.frame = (long)frame;
                const long units = (long)n * x->A * g.row16;
                const int grid = grid_for((units + 1023) / 1024, device_sm_count(x->device), 4);
                k_xe_ingest<<<grid, 256, 0, st.s_h2d>>>(g);
                CLB_CUDA(cudaGetLastError());
                x->n_launch++;
            } else {
                for (int s = 0; s < x->A; s++)
                    CLB_CUDA(cudaMemcpy2DAsync(dst_dev + s * row, frame, (const char *)ports[s] + (size_t)done * vec_in + off_in,
                                               vec_in, row, (size_t)n, cudaMemcpyHostToDevice, st.s_h2d));
            }
        } else {
            char *stage = (char *)st.pin[b].p + (size_t)st.tracker * frame;
            for (long t = 0; t < n; t++) {
                char *d = stage + (size_t)t * frame;
                for (int s = 0; s < x->A; s++) {
                    if (!planar_pol) {
                        memcpy(d + s * row, (const char *)ports[s] + (size_t)(done + t) * vec_in + off_in, row);
                    } else {
                        const char *px = (const char *)ports[s] + (size_t)(done + t) * vec_in + off_in;
                        const char *py = (const char *)ports[s + x->A] + (size_t)(done + t) * vec_in + off_in;
                        char *o = d + s * row;
                        if (sb == 2) {
                            for (int c = 0; c < x->F; c++) {
                                memcpy(o + 4 * c, px + 2 * c, 2);
                                memcpy(o + 4 * c + 2, py + 2 * c, 2);
                            }
                        } else {
                            for (int c = 0; c < x->F; c++) {
                                memcpy(o + 2 * c * sb, px + c * sb, sb);
                                memcpy(o + (2 * c + 1) * sb, py + c * sb, sb);
                            }
                        }
                    }
                }
            }
            CLB_CUDA(cudaMemcpyAsync(dst_dev, stage, (size_t)n * frame, cudaMemcpyHostToDevice, st.s_h2d));
            st.pin_dirty[b] = true;
        }
        x->n_h2d += (size_t)n * frame;
        st.tracker += n;
        done += n;
        // the ports are the caller's again on return -- unless the caller has promised to leave them alone
        // (clb200_xengine_stream_ports_stable): then the DMA of this push overlaps the caller's next one
        if (direct && !st.ports_stable) CLB_CUDA(cudaStreamSynchronize(st.s_h2d));
        if (st.tracker == x->T) {
            CLB_TRY(xe_stream_launch(x, b));
            st.tracker = 0;
        }
    }
    return CLB200_OK;
}

int xe_stream_poll(XEngine *x, void *out_c32, int wait, int *ready)
{
    XEngine::Stream &st = x->st;
    CLB_CHECK(st.on, CLB200_ESTATE, "clXEngine: stream_begin first");
    *ready = 0;
    if (st.res_tail >= st.res_head) return CLB200_OK;
    const int r = (int)(st.res_tail % st.nres);
    if (wait) {
        CLB_CUDA(cudaEventSynchronize(st.ev_res[r]));
    } else {
        cudaError_t e = cudaEventQuery(st.ev_res[r]);
        if (e == cudaErrorNotReady) return CLB200_OK;
        CLB_CUDA(e);
    }
    if (out_c32) memcpy(out_c32, st.pin_res[r].p, (size_t)x->out_items() * 8);
    st.res_tail++;
    *ready = 1;
    return CLB200_OK;
}

} // namespace

extern "C" {

int clb200_xengine_create(int device, int data_type, int npol, int num_inputs, int num_channels,
                          int integration, clb200_handle *out)
{
    CLB_CHECK(out != nullptr, CLB200_EINVAL, "null out");
    // lib/clXEngine_impl.cc:106-109 throws std::out_of_range for fewer than 2 inputs
    CLB_CHECK(num_inputs >= 2, CLB200_EINVAL, "clXEngine: Please specify at least 2 inputs (got %d)", num_inputs);
    CLB_CHECK(npol == 1 || npol == 2, CLB200_EINVAL, "clXEngine: polarization must be 1 or 2, got %d", npol);
    CLB_CHECK(data_type == CLB200_DTYPE_COMPLEX || data_type == CLB200_DTYPE_BYTE ||
                  data_type == CLB200_DTYPE_PACKEDXY,
              CLB200_EINVAL, "clXEngine: data type %d is not complex(1), byte(5) or packed-XY(6)", data_type);
    CLB_CHECK(num_channels >= 1 && integration >= 1, CLB200_EINVAL,
              "clXEngine: num_channels and integration must be positive");
    CLB_CHECK((long)num_inputs * num_channels * npol * 2 * 32 < (1L << 31), CLB200_EINVAL,
              "clXEngine: one time step (%d inputs x %d channels) is too large for the stage addressing",
              num_inputs, num_channels);
    CLB_CHECK(((long)num_channels + 3) / 4 * (((long)integration + 31) / 32) < (1L << 31), CLB200_EINVAL,
              "clXEngine: %d channels x %d time steps exceed the stage index range", num_channels, integration);
    CLB_CHECK(num_inputs * npol <= 64, CLB200_EINVAL,
              "clXEngine: %d inputs x %d polarisations exceeds the 64 rows this build tiles for",
              num_inputs, npol);
    int n = clb200_device_count();
    CLB_CHECK(n > 0, CLB200_ECUDA, "no CUDA device present");
    CLB_CHECK(device >= 0 && device < n, CLB200_EINVAL, "device %d out of range", device);
    DeviceGuard g(device);
    XEngine *x = new XEngine;
    x->kind = KIND_XENGINE;
    x->device = device;
    x->data_type = data_type;
    x->npol = npol;
    x->A = num_inputs;
    x->F = num_channels;
    x->T = integration;
    x->var = pick_xe(num_inputs * npol);
    cudaError_t e = cudaFuncSetAttribute((const void *)x->var->kernel[npol - 1],
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         x->var->smem_bytes);
    if (e == cudaSuccess && num_inputs * npol <= 32) {
        const char *legacy = getenv("CLB200_XE_LEGACY");       // force the mma.sync kernel (A/B testing)
        x->use_tc = !(legacy && atoi(legacy));
        if (x->use_tc)
            e = npol == 1 ? cudaFuncSetAttribute((const void *)k_xengine_tc<1>,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM)
                          : cudaFuncSetAttribute((const void *)k_xengine_tc<2>,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM);
        const char *notma = getenv("CLB200_XE_TMA");           // 0: keep the LDG-fed tcgen05 kernel
        if (x->use_tc && e == cudaSuccess && !(notma && atoi(notma) == 0)) {
            cudaError_t e2 = cudaFuncSetAttribute((const void *)tm_kernel(npol, 8),
                                                  cudaFuncAttributeMaxDynamicSharedMemorySize, TM_SMEM);
            if (e2 == cudaSuccess)
                e2 = cudaFuncSetAttribute((const void *)tm_kernel(npol, 16),
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, TM_SMEM);
            x->use_tma = (e2 == cudaSuccess) && tm_encoder() != nullptr;
            const char *up = getenv("CLB200_XE_UNPACK_PASS");       // 1: keep the separate unpack pass (A/B measurements)
            x->use_tma_pk = x->use_tma && !(up && atoi(up)) && cudaFuncSetAttribute((const void *)tm_kernel(npol, 16, true),
                                                               cudaFuncAttributeMaxDynamicSharedMemorySize, TM_SMEM) == cudaSuccess;
            const char *fo = getenv("CLB200_XE_FC");
            x->fc_override = fo ? atoi(fo) : 0;
            const char *pd = getenv("CLB200_XE_PDL");
            x->pdl = !(pd && atoi(pd) == 0);
            const char *pr = getenv("CLB200_XE_L2PROMO");
            x->l2promo = pr ? atoi(pr) : 0;
        }
    }
    if (e != cudaSuccess) {
        set_error("clXEngine: cannot reserve %d B of shared memory: %s", x->var->smem_bytes,
                  cudaGetErrorString(e));
        delete x;
        return CLB200_ECUDA;
    }
    x->set_info("clXEngine %d inputs x %d pol, %d channels, %d time steps, data type %d: %s", num_inputs, npol, num_channels,
                integration, data_type,
                data_type == CLB200_DTYPE_COMPLEX ? "tiled FP32 kernel (complex float input)"
                : x->use_tma ? "k_xengine_tma: TMA boxes -> LDSM/STSM byte transposes -> tcgen05.mma kind::i8 (s32 in TMEM), "
                               "16 (or 8) channels per CTA, time slices as thread-block clusters"
                : x->use_tc ? "k_xengine_tc: LDG-fed tcgen05.mma kind::i8"
                            : "k_xengine_i8: mma.sync m16n8k32 (33..64 input rows)");
    *out = x;
    return CLB200_OK;
}

long clb200_xengine_input_bytes(clb200_handle h)
{
    XEngine *x;
    if (check_kind(h, KIND_XENGINE, &x) != CLB200_OK) return CLB200_EINVAL;
    const int Ftot = x->Ftotal > 0 ? x->Ftotal : x->F;
    return (long)x->T * x->A * Ftot * x->npol * (long)x->sample_bytes();
}

long clb200_xengine_output_items(clb200_handle h)
{
    XEngine *x;
    if (check_kind(h, KIND_XENGINE, &x) != CLB200_OK) return CLB200_EINVAL;
    return x->out_items();
}

int clb200_xengine_set_shard(clb200_handle h, int total_channels, int chan_first)
{
    XEngine *x;
    CLB_TRY(check_kind(h, KIND_XENGINE, &x));
    CLB_CHECK(chan_first >= 0 && chan_first + x->F <= total_channels, CLB200_EINVAL,
              "clXEngine: shard [%d, %d) does not fit %d channels", chan_first, chan_first + x->F,
              total_channels);
    x->Ftotal = total_channels;
    x->f_first = chan_first;
    return CLB200_OK;
}

int clb200_xengine_set_gather(clb200_handle h, int nranks, void *const *full_out_c32)
{
    XEngine *x;
    CLB_TRY(check_kind(h, KIND_XENGINE, &x));
    CLB_CHECK(nranks >= 1 && nranks <= CLB200_XENGINE_MAX_GATHER && full_out_c32, CLB200_EINVAL,
              "clXEngine: 1..%d gather destinations, got %d", CLB200_XENGINE_MAX_GATHER, nranks);
    CLB_CHECK(x->Ftotal > 0, CLB200_ESTATE, "clXEngine: set_shard first");
    CLB_CHECK(x->data_type != CLB200_DTYPE_COMPLEX, CLB200_EINVAL, "clXEngine: the gather is for the integer input types");
    for (int r = 0; r < nranks; r++) {
        CLB_CHECK(full_out_c32[r] != nullptr && ((uintptr_t)full_out_c32[r] % 8) == 0, CLB200_EINVAL, "bad destination %d", r);
        x->gather[r] = full_out_c32[r];
    }
    x->ngather = nranks;
    return CLB200_OK;
}

int clb200_xengine_set_gather_sync(clb200_handle h, int my_rank, void *const *flag_arrays, void *multicast_out,
                                   void *multicast_flags)
{
    XEngine *x;
    CLB_TRY(check_kind(h, KIND_XENGINE, &x));
    CLB_CHECK(x->ngather > 0, CLB200_ESTATE, "clXEngine: set_gather first");
    CLB_CHECK(my_rank >= 0 && my_rank < x->ngather && flag_arrays != nullptr, CLB200_EINVAL, "bad rank / flag arrays");
    CLB_CHECK(((uintptr_t)multicast_out % 16) == 0, CLB200_EINVAL, "multicast address must be 16 B aligned");
    DeviceGuard g(x->device);
    for (int r = 0; r < x->ngather; r++) {
        CLB_CHECK(flag_arrays[r] != nullptr, CLB200_EINVAL, "null flag array %d", r);
        x->gather_flags[r] = (unsigned *)flag_arrays[r];
    }
    CLB_TRY(x->d_gather_counter.reserve(256));
    CLB_CUDA(cudaMemset(x->d_gather_counter.p, 0, 256));
    x->gather_rank = my_rank;
    x->gather_mc = multicast_out;
    x->gather_flags_mc = (unsigned *)multicast_flags;
    x->gather_epoch = 0;
    return CLB200_OK;
}

int clb200_xengine_gather_wait(clb200_handle h, void *stream)
{
    XEngine *x;
    CLB_TRY(check_kind(h, KIND_XENGINE, &x));
    CLB_CHECK(x->gather_rank >= 0, CLB200_ESTATE, "clXEngine: set_gather_sync first");
    DeviceGuard g(x->device);
    XeGatherSync gs;
    for (int r = 0; r < 8; r++)
        gs.flag[r] = (r < x->ngather && x->gather_flags[r]) ? x->gather_flags[r] + (size_t)x->gather_rank * CLB200_XENGINE_FLAG_STRIDE : nullptr;
    gs.flag_mc = x->gather_flags_mc ? x->gather_flags_mc + (size_t)x->gather_rank * CLB200_XENGINE_FLAG_STRIDE : nullptr;
    gs.local = x->gather_flags[x->gather_rank];
    gs.epoch = x->gather_epoch;
    gs.nranks = x->ngather;
    {
        const char *e = getenv("CLB200_XE_GATHER_INKERNEL");
        gs.signal = !(e && atoi(e));             // the kernel's last CTA has released the flags already
    }
    gs.wait = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(1);
    cfg.blockDim = dim3(32);
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = x->pdl ? 1 : 0;
    CLB_CUDA(cudaLaunchKernelEx(&cfg, k_gather_signal_wait, gs));
    x->n_launch++;
    return CLB200_OK;
}

int clb200_xengine_launch_device_gather(clb200_handle h, const void *d_in, void *stream)
{
    XEngine *x;
    CLB_TRY(check_kind(h, KIND_XENGINE, &x));
    CLB_CHECK(x->ngather > 0, CLB200_ESTATE, "clXEngine: set_gather first");
    DeviceGuard g(x->device);
    // d_in: this rank's channel slab only, [t][station][shard channels][pol] (what work() uploads per rank)
    return xe_launch(x, d_in, x->T, x->F, 0, nullptr, nullptr, 0, (cudaStream_t)stream, true);
}

int clb200_xengine_work(clb200_handle h, const void *in, void *out_c32, int accumulate)
{
    XEngine *x;
    CLB_TRY(check_kind(h, KIND_XENGINE, &x));
    CLB_CHECK(in && out_c32, CLB200_EINVAL, "null buffer");
    DeviceGuard g(x->device);
    return xe_work(x, in, out_c32, false, accumulate);
}

int clb200_xengine_work_i32(clb200_handle h, const void *in, int32_t *out_i32)
{
    XEngine *x;
    CLB_TRY(check_kind(h, KIND_XENGINE, &x));
    CLB_CHECK(in && out_i32, CLB200_EINVAL, "null buffer");
    DeviceGuard g(x->device);
    return xe_work(x, in, out_i32, true, 0);
}

int clb200_xengine_launch_device(clb200_handle h, const void *d_in, void *d_out_c32, int accumulate,
                                 void *stream)
{
    XEngine *x;
    CLB_TRY(check_kind(h, KIND_XENGINE, &x));
    DeviceGuard g(x->device);
    const int Ftot = x->Ftotal > 0 ? x->Ftotal : x->F;
    return xe_launch(x, d_in, x->T, Ftot, x->f_first, nullptr, (float2 *)d_out_c32, accumulate,
                     (cudaStream_t)stream);
}

int clb200_xengine_launch_device_batch(clb200_handle h, const void *d_in, void *d_out_c32, int nbatch, void *stream)
{
    XEngine *x;
    CLB_TRY(check_kind(h, KIND_XENGINE, &x));
    CLB_CHECK(nbatch >= 1 && nbatch <= 4096, CLB200_EINVAL, "clXEngine: batch of %d integrations", nbatch);
    CLB_CHECK(x->data_type == CLB200_DTYPE_BYTE, CLB200_EINVAL, "clXEngine: batched launches take IChar input");
    DeviceGuard g(x->device);
    const int Ftot = x->Ftotal > 0 ? x->Ftotal : x->F;
    const size_t in_stride = (size_t)x->T * x->A * Ftot * x->npol * 2;
    const long rowb = (long)Ftot * x->npol * 2;
    // one grid over all integrations with the TMA kernel (one channel group, or one time slice of it, per CTA);
    // else one launch each
    const bool one_grid = x->use_tc && x->use_tma && ((uintptr_t)d_in % 16 == 0) && (rowb % 16 == 0) &&
                          ((uintptr_t)d_out_c32 % 8 == 0) && nbatch > 1;
    if (one_grid)
        return xe_launch(x, d_in, x->T, Ftot, x->f_first, nullptr, (float2 *)d_out_c32, 0, (cudaStream_t)stream, false, nbatch);
    for (int k = 0; k < nbatch; k++)
        CLB_TRY(xe_launch(x, (const char *)d_in + (size_t)k * in_stride, x->T, Ftot, x->f_first, nullptr,
                          (float2 *)d_out_c32 + (size_t)k * x->out_items(), 0, (cudaStream_t)stream));
    return CLB200_OK;
}

int clb200_xengine_launch_device_i32(clb200_handle h, const void *d_in, int32_t *d_out_i32, void *stream)
{
    XEngine *x;
    CLB_TRY(check_kind(h, KIND_XENGINE, &x));
    CLB_CHECK(x->data_type != CLB200_DTYPE_COMPLEX, CLB200_EINVAL,
              "clXEngine: complex input has no integer output");
    DeviceGuard g(x->device);
    const int Ftot = x->Ftotal > 0 ? x->Ftotal : x->F;
    return xe_launch(x, d_in, x->T, Ftot, x->f_first, d_out_i32, nullptr, 0, (cudaStream_t)stream);
}

int clb200_xengine_stream_begin(clb200_handle h, int pipeline_integration, int result_slots)
{
    XEngine *x;
    CLB_TRY(check_kind(h, KIND_XENGINE, &x));
    DeviceGuard g(x->device);
    std::lock_guard<std::mutex> lk(x->mtx);
    int rc = xe_stream_begin(x, pipeline_integration, result_slots > 0 ? result_slots : 4);
    if (rc != CLB200_OK) x->stream_free();
    return rc;
}

int clb200_xengine_push_timesteps(clb200_handle h, const void *const *ports, int nports, long ntime)
{
    XEngine *x;
    CLB_TRY(check_kind(h, KIND_XENGINE, &x));
    CLB_CHECK(ports != nullptr && ntime >= 0, CLB200_EINVAL, "bad arguments");
    for (int i = 0; i < nports; i++) CLB_CHECK(ports[i] != nullptr, CLB200_EINVAL, "null port %d", i);
    DeviceGuard g(x->device);
    std::lock_guard<std::mutex> lk(x->mtx);
    return xe_stream_push(x, ports, nports, ntime);
}

int clb200_xengine_poll_result(clb200_handle h, void *out_c32, int wait, int *ready)
{
    XEngine *x;
    CLB_TRY(check_kind(h, KIND_XENGINE, &x));
    CLB_CHECK(ready != nullptr, CLB200_EINVAL, "null ready");
    DeviceGuard g(x->device);
    std::lock_guard<std::mutex> lk(x->mtx);
    return xe_stream_poll(x, out_c32, wait, ready);
}

int clb200_xengine_stream_state(clb200_handle h, long *tracker, long *integrations, long *results_pending,
                                uint64_t *pushes, uint64_t *pushes_blocked)
{
    XEngine *x;
    CLB_TRY(check_kind(h, KIND_XENGINE, &x));
    std::lock_guard<std::mutex> lk(x->mtx);
    if (tracker) *tracker = x->st.tracker;
    if (integrations) *integrations = x->st.n_integ;
    if (results_pending) *results_pending = x->st.res_head - x->st.res_tail;
    if (pushes) *pushes = x->st.n_push;
    if (pushes_blocked) *pushes_blocked = x->st.n_blocked;
    return CLB200_OK;
}

int clb200_xengine_stream_ports_stable(clb200_handle h, int stable)
{
    XEngine *x;
    CLB_TRY(check_kind(h, KIND_XENGINE, &x));
    std::lock_guard<std::mutex> lk(x->mtx);
    CLB_CHECK(x->st.on, CLB200_ESTATE, "clXEngine: stream_begin first");
    x->st.ports_stable = stable != 0;
    return CLB200_OK;
}

int clb200_xengine_stream_end(clb200_handle h)
{
    XEngine *x;
    CLB_TRY(check_kind(h, KIND_XENGINE, &x));
    DeviceGuard g(x->device);
    std::lock_guard<std::mutex> lk(x->mtx);
    x->stream_free();
    return CLB200_OK;
}

} // extern "C"
