// xengine_tma.cuh -- clXEngine, tcgen05/TMEM contraction fed by TMA + hardware byte transposes.
//
// Same arithmetic, TMEM layout, MMA issue and epilogue as k_xengine_tc (xengine_tc.cuh):
// per channel G = Z Z^T with Z[(input,re|im)][t] int8, exact in s32 (lib/clXEngine_impl.cc:729-810
// restated on integers).  What changes is how Z gets into shared memory:
//   * one producer lane issues cp.async.bulk.tensor loads of raw [station][32 t][FC channels]
//     boxes (32 KiB) into a 4-deep ring, tracked by tx-count mbarriers -- 128 KiB per SM
//     in flight instead of the two register stages of the LDG feed, which is what an
//     HBM-latency-bound kernel needs;
//   * the [t][station][chan] wire layout has time outermost and the tensor cores want it
//     innermost.  The box is written station-major with the TMA 32 B / 64 B swizzle, so that
//     `ldmatrix.m16n16.trans.b8` (LDSM.8.MT1616) reads 16 time steps x 16 B of one station
//     bank-conflict free and hands every thread 4 consecutive time steps of one
//     (channel, re|im) byte lane; `stmatrix.m8n8.x4` then writes 16 B rows (16 time steps)
//     of the UMMA canonical K-major image, 8 different channel images per phase (the
//     per-channel skew keeps those on 8 different bank groups).  Two instructions move
//     512 B; the PRMT transposes and per-word address arithmetic of the LDG feed are gone.
// Requirements (checked on the host, else k_xengine_tc runs): 16 B aligned input rows.
#pragma once

#include <cuda.h>

namespace {

constexpr int TM_NR = 3;                       // raw ring depth
constexpr int TM_NI = 3;                       // image ring depth
constexpr int TM_RAWB = 32 * 1024;             // bytes per raw box: 32 t x (32 stations x 32 B | 16 x 64 B)
constexpr int TM_IMGB = 16 * (2048 + 64);      // bytes reserved per stage image (FC channels x KT steps x 64 rows + skew)
constexpr int TM_SMEM = 1024 + TM_NR * TM_RAWB + TM_NI * TM_IMGB + 256;
constexpr int TM_THREADS = XE_THREADS + 64;    // 16 transpose/epilogue warps + MMA warp + TMA warp

__device__ __forceinline__ void tm_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tm_load_3d(uint32_t dst, const CUtensorMap *tm, uint64_t *bar, int c0, int c1, int c2)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(tm), "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tm_ldsm_t8(uint32_t addr, uint32_t (&r)[4])
{
    asm volatile("ldmatrix.sync.aligned.m16n16.x2.trans.shared.b8 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr)
                 : "memory");
}
__device__ __forceinline__ void tm_stsm(uint32_t addr, const uint32_t (&r)[4])
{
    asm volatile("stmatrix.sync.aligned.m8n8.x4.shared.b16 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3])
                 : "memory");
}
// the 16 transpose/epilogue warps only (the MMA and TMA warps never join)
__device__ __forceinline__ void tm_worker_sync() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

// Extra launch state of the TMA kernel: where partial sums of time-sliced CTAs meet.
struct TmFix {
    int2 *part;             // [nslice][F][nbl*npol^2] partial visibilities (re, im) of each time slice
    unsigned *count;        // [groups] arrival counters, zero between launches (the last CTA resets them)
    long long *stamp;       // debug time stamps (or null)
};

// phase time stamps of CTA 0 (CLB200_XE_DBG & 8): fx.stamp[k] = clock64 at phase k
#define TM_STAMP(k)                                                                                  \
    do {                                                                                             \
        if ((p.dbg & 8) && blockIdx.x == 0 && threadIdx.x == 0 && fx.stamp) fx.stamp[k] = clock64(); \
    } while (0)

// FC = channels per CTA: 16 fills the 512 TMEM columns; 8 (256 columns) lets 1024 channels spread over
// 128 CTAs WITHOUT slicing time, i.e. without any cross-CTA reduction.  A raw box is always 32 KiB:
// KT = 512 / FC time steps (KS = KT / 32 MMA k-steps per stage).
template <int NPOL, int FC>
__global__ void __launch_bounds__(TM_THREADS, 1)
k_xengine_tma(XeParams p, TmFix fx, const __grid_constant__ CUtensorMap tmap)
{
    constexpr int IB = FC * NPOL * 2;                        // bytes per (t, station) run: 16 | 32 | 64
    constexpr int NH = IB / 16;                              // 16 B chunks per run
    constexpr int KT = 512 / FC;                             // time steps per stage
    constexpr int KS = KT / 32;                              // MMA k-steps per stage
    constexpr int ASTN = 32 / NPOL;                          // stations per box
    constexpr int CSK = KS * 2048 + ((NPOL == 1) ? 32 : 64); // channel image stride: the skew spreads the STSM banks
    constexpr int UPW = 64 / XE_WARPS;                       // ldmatrix.x2 units per warp and stage
    static_assert(FC * CSK <= TM_IMGB, "stage image does not fit");
    static_assert(ASTN * KT * IB == TM_RAWB, "raw box is 32 KiB");

    extern __shared__ uint8_t tm_smem_raw[];
    // the swizzle pattern is a function of the shared-memory ADDRESS: align the ring to 1 KiB
    const uint32_t sbase = ((uint32_t)__cvta_generic_to_shared(tm_smem_raw) + 1023u) & ~1023u;
    uint8_t *sptr = tm_smem_raw + (sbase - (uint32_t)__cvta_generic_to_shared(tm_smem_raw));
    const uint32_t raw_addr = sbase;
    const uint32_t img_addr = sbase + TM_NR * TM_RAWB;
    int2 *stg = reinterpret_cast<int2 *>(sptr + TM_NR * TM_RAWB);            // epilogue staging = image ring
    uint64_t *bars = reinterpret_cast<uint64_t *>(sptr + TM_NR * TM_RAWB + TM_NI * TM_IMGB);
    uint64_t *full = bars;                           // [NI]  stage image complete
    uint64_t *empty = bars + TM_NI;                  // [NI]  MMAs that read the image are done
    uint64_t *rfull = bars + 2 * TM_NI;              // [NR]  raw box landed
    uint64_t *rempty = bars + 2 * TM_NI + TM_NR;     // [NR]  raw box transposed
    uint64_t *tfree = bars + 2 * TM_NI + 2 * TM_NR;  //       epilogue drained TMEM
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tfree + 1);
    unsigned *last_flag = tmem_slot + 1;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int CS = (p.dbg & 16) ? KS * 2048 : CSK;            // experiment: unskewed (128 B aligned) images
    const int nbl = p.A * (p.A + 1) / 2;
    const int npp = nbl * NPOL * NPOL;                       // visibilities per channel
    const int spp = npp + 4;                                 // staging stride per channel (int2): breaks the bank tie
    const int ngroups = (p.F + FC - 1) / FC;
    const int nst = (p.T + KT - 1) / KT;

    int s0, s1, slice = 0;
    if (p.nslice > 0) {
        const int grp = blockIdx.x % ngroups;
        slice = blockIdx.x / ngroups;
        const int len = (nst + p.nslice - 1) / p.nslice;
        s0 = grp * nst + slice * len;
        s1 = min(grp * nst + nst, s0 + len);
    } else if (p.split) {
        const long total = (long)ngroups * nst;
        s0 = (int)(total * blockIdx.x / gridDim.x);
        s1 = (int)(total * (blockIdx.x + 1) / gridDim.x);
    } else {
        s0 = (int)((long)ngroups * blockIdx.x / gridDim.x) * nst;
        s1 = (int)((long)ngroups * (blockIdx.x + 1) / gridDim.x) * nst;
    }
    // a time slice may be empty (more slices than stages); it still takes part in the slice fix-up
    const int nstages = max(0, s1 - s0);
    if (nstages == 0 && !(p.nslice > 1)) return;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
            (uint32_t)__cvta_generic_to_shared(tmem_slot)), "n"(FC * 32));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < TM_NI; i++) {
            tc_mbar_init(&full[i], XE_WARPS);
            tc_mbar_init(&empty[i], 1);
        }
        for (int r = 0; r < TM_NR; r++) {
            tc_mbar_init(&rfull[r], 1);
            tc_mbar_init(&rempty[r], XE_WARPS);
        }
        tc_mbar_init(tfree, XE_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == XE_WARPS + 1) {
        // ---- TMA producer ----
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
            int grp = s0 / nst, st = s0 - grp * nst;
            for (int n = 0; n < nstages; n++) {
                const int r = n % TM_NR;
                if (n >= TM_NR) tc_mbar_wait(&rempty[r], (uint32_t)(n / TM_NR - 1) & 1u);
                tm_expect_tx(&rfull[r], TM_RAWB);
                tm_load_3d(raw_addr + r * TM_RAWB, &tmap, &rfull[r], (p.f_off + grp * FC) * NPOL * 2, st * KT, 0);
                if (++st == nst) {
                    st = 0;
                    grp++;
                }
            }
        }
    } else if (warp == XE_WARPS) {
        // ---- MMA issue: the warp stays converged, one ELECTED lane issues (a plain `lane == 0` branch
        // makes the compiler wrap every tcgen05.mma in a per-active-lane loop: ~12 instructions each) ----
        uint32_t leader;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
        int groups_done = 0;
        for (int n = 0; n < nstages; n++) {
            const int sg = s0 + n, b = n % TM_NI;
            const bool group_first = (n == 0) || (sg % nst == 0);
            tc_mbar_wait(&full[b], (uint32_t)(n / TM_NI) & 1u);
            if (group_first && n > 0) tc_mbar_wait(tfree, (uint32_t)(groups_done - 1) & 1u);
            tc_fence_after();
            if (leader) {
                if (!(p.dbg & 1))
#pragma unroll
                    for (int ks = 0; ks < KS; ks++)
#pragma unroll
                        for (int ch = 0; ch < FC; ch++) {
                            const uint64_t d = tc_smem_desc(img_addr + b * TM_IMGB + ch * CS + ks * 2048);
                            tc_mma_i8(tmem_base + (((uint32_t)(ch & 1) * 16u) << 16) + (uint32_t)(ch >> 1) * 64u, d, d,
                                      (group_first && ks == 0) ? 0u : 1u);
                        }
                tc_commit(&empty[b]);
            }
            __syncwarp();
            if ((n + 1 == nstages) || ((sg + 1) % nst == 0)) groups_done++;
        }
    } else {
        // ---- transpose (raw box -> K-major images) + epilogue ----
        // unit u of a stage = what one ldmatrix.x2 covers: station s and either two 16 B chunks of one
        // 16-step k chunk (runs of 32 / 64 B) or the single chunk of two k chunks (16 B runs)
        uint32_t ld_off[UPW], st_off[UPW];
#pragma unroll
        for (int i = 0; i < UPW; i++) {
            const int u = warp * UPW + i;
            constexpr int UPS = 64 / ASTN;                       // units per station
            const int s = u / UPS, w = u % UPS;
            const int kc0 = (NH == 1) ? 2 * w : w / (NH / 2 > 0 ? NH / 2 : 1);
            const int h0 = (NH == 1) ? 0 : 2 * (w % (NH / 2 > 0 ? NH / 2 : 1));
            {   // ldmatrix: lanes 0..15 address the 16 rows (time steps) of matrix 0, lanes 16..31 of matrix 1
                const int mi = lane >> 4;
                const int kc = (NH == 1) ? kc0 + mi : kc0, h = (NH == 1) ? 0 : h0 + mi;
                const int t = kc * 16 + (lane & 15);
                const int sw = (NH == 1) ? 0 : (NH == 2) ? ((t >> 2) & 1) : ((t >> 1) & 3);
                ld_off[i] = (uint32_t)(s * (KT * IB) + t * IB + ((h ^ sw) << 4));
            }
            {   // stmatrix: lane L addresses row (L & 7) of matrix (L >> 3) = register (L >> 3);
                // registers 2*mi + j hold byte lanes j*8 .. j*8+7 of ldmatrix matrix mi
                const int q = lane >> 3, rr = lane & 7, mi = q >> 1;
                const int kc = (NH == 1) ? kc0 + mi : kc0, h = (NH == 1) ? 0 : h0 + mi;
                const int B = h * 16 + (q & 1) * 8 + rr;                 // byte of the (t, station) run
                const int f = (NPOL == 1) ? (B >> 1) : (B >> 2);
                const int v = (NPOL == 1) ? s : 2 * s + ((B >> 1) & 1);
                const int m = 2 * v + (B & 1);
                st_off[i] = (uint32_t)(f * CS + (kc >> 1) * 2048 + (m >> 3) * 256 + (kc & 1) * 128 + (m & 7) * 16);
            }
        }
        const int tid = threadIdx.x;                               // 0..511 (worker warps come first)
        TM_STAMP(4);
        for (int n = 0; n < max(nstages, 1); n++) {
            const int sg = s0 + n, b = n % TM_NI, r = n % TM_NR;
            if (nstages > 0) {
                tc_mbar_wait(&rfull[r], (uint32_t)(n / TM_NR) & 1u);
                if (n >= TM_NI) tc_mbar_wait(&empty[b], (uint32_t)(n / TM_NI - 1) & 1u);   // MMAs of stage n-NI are done
                const uint32_t src = raw_addr + r * TM_RAWB, dst = img_addr + b * TM_IMGB;
                if (!(p.dbg & 2)) {
                    uint32_t v[UPW][4];
#pragma unroll
                    for (int i = 0; i < UPW; i++) tm_ldsm_t8(src + ld_off[i], v[i]);
#pragma unroll
                    for (int i = 0; i < UPW; i++) tm_stsm(dst + st_off[i], v[i]);
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tc_mbar_arrive(&full[b]);
                    tc_mbar_arrive(&rempty[r]);
                }
                const bool group_done = (n + 1 == nstages) || ((sg + 1) % nst == 0);
                if (!group_done) continue;
                tc_mbar_wait(&empty[b], (uint32_t)(n / TM_NI) & 1u);       // every MMA of the group has landed
                tc_fence_after();
            }
            if (p.dbg & 4) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) tc_mbar_arrive(tfree);
                continue;
            }
            // ---- epilogue.  All MMAs of the group are done and no worker is ahead of this point, so the
            // image ring is free: the 16 channels are combined (re/im rows sit in adjacent lanes), staged
            // there in OUTPUT order [channel][baseline][pol^2](re, im) and leave as ONE contiguous block
            // (the output is channel-major, lib/clXEngine_impl.cc:799-806).
            TM_STAMP(0);
            const int grp_e = (p.nslice > 0) ? (int)(blockIdx.x % ngroups) : sg / nst;
            const int f0 = grp_e * FC;
            if (nstages > 0) {
                const int wq = warp & 3;
                const int v1 = 8 * wq + ((lane & 15) >> 1), c1 = lane & 1;
                const int st1 = (NPOL == 1) ? v1 : (v1 >> 1);
                const bool row_ok = st1 < p.A;
                int *stw = reinterpret_cast<int *>(stg);
#pragma unroll 1
                for (int jj = 0; jj < FC / 8; jj++) {
                    const int chl = 2 * ((warp >> 2) + 4 * jj) + (lane >> 4);
                    int ob;                                               // staging index for column input 0
                    if (NPOL == 1) ob = chl * spp + st1 * (st1 + 1) / 2;
                    else ob = chl * spp + 4 * (st1 * (st1 + 1) / 2) + 2 * (v1 & 1);
#pragma unroll 1
                    for (int half = 0; half < 2; half++) {
                        uint32_t rg[32];
                        tc_ld32(tmem_base + (((uint32_t)wq * 32u) << 16) +
                                    (uint32_t)(((warp >> 2) + 4 * jj) * 64 + half * 32), rg);
#pragma unroll
                        for (int i = 0; i < 16; i++) {
                            const int v2 = 16 * half + i;
                            const int mine = (int)rg[2 * i];
                            const int other = __shfl_xor_sync(0xffffffffu, (int)rg[2 * i + 1], 1);
                            const int val = c1 ? mine - other : mine + other;
                            const int st2 = (NPOL == 1) ? v2 : (v2 >> 1);
                            const int o = ob + ((NPOL == 1) ? v2 : (4 * st2 + (v2 & 1)));
                            if (row_ok && st2 <= st1) stw[2 * o + c1] = val;
                        }
                    }
                }
                tc_fence_before();                                     // TMEM is drained
                __syncwarp();
                if (lane == 0) tc_mbar_arrive(tfree);
            }
            tm_worker_sync();
            TM_STAMP(1);
            const int nch = min(FC, p.F - f0);
            const int nit = nch * npp;                                 // int2 items of this block
            const long gbase = (long)f0 * npp;
            const bool fix = p.nslice > 1;
            bool last = true;
            constexpr int UN = 4;
            if (fix) {
                // every slice publishes its partial block; the last one to arrive folds the others in
                if (nstages > 0) {
                    int2 *dst = fx.part + (long)slice * p.F * npp + gbase;
                    for (int i = tid; i < nit; i += XE_THREADS) {
                        const int ch = i / npp;
                        dst[i] = stg[i + ch * (spp - npp)];
                    }
                }
                __threadfence();
                tm_worker_sync();
                if (tid == 0) *last_flag = atomicAdd(&fx.count[grp_e], 1u);
                tm_worker_sync();
                last = (*last_flag == (unsigned)(p.nslice - 1));
                if (last) __threadfence();
            }
            TM_STAMP(2);
            if (last) {
                const int len = (nst + p.nslice - 1) / max(p.nslice, 1);
                for (int i0 = tid; i0 < nit; i0 += UN * XE_THREADS) {
                    int2 v[UN];
                    int2 oi[UN];
                    float2 of[UN];
#pragma unroll
                    for (int k = 0; k < UN; k++) {
                        const int i = i0 + k * XE_THREADS;
                        v[k] = make_int2(0, 0);
                        if (i < nit && nstages > 0) v[k] = stg[i + (i / npp) * (spp - npp)];
                    }
                    if (fix)
                        for (int s = 0; s < p.nslice; s++) {
                            if (s == slice || s * len >= nst) continue;         // own or empty slice
                            const int2 *src = fx.part + (long)s * p.F * npp + gbase;
                            int2 w[UN];
#pragma unroll
                            for (int k = 0; k < UN; k++) {
                                const int i = i0 + k * XE_THREADS;
                                w[k] = (i < nit) ? __ldcg(src + i) : make_int2(0, 0);
                            }
#pragma unroll
                            for (int k = 0; k < UN; k++) {
                                v[k].x += w[k].x;
                                v[k].y += w[k].y;
                            }
                        }
                    if (p.split && !fix) {                         // stream-K: partial groups meet in memory
#pragma unroll
                        for (int k = 0; k < UN; k++) {
                            const int i = i0 + k * XE_THREADS;
                            if (i < nit) {
                                atomicAdd(p.out_i32 + 2 * (gbase + i), v[k].x);
                                atomicAdd(p.out_i32 + 2 * (gbase + i) + 1, v[k].y);
                            }
                        }
                        continue;
                    }
                    if (p.accumulate) {
#pragma unroll
                        for (int k = 0; k < UN; k++) {
                            const int i = i0 + k * XE_THREADS;
                            if (i < nit && p.out_i32) oi[k] = reinterpret_cast<int2 *>(p.out_i32)[gbase + i];
                            if (i < nit && p.out_f32) of[k] = p.out_f32[gbase + i];
                        }
                    }
#pragma unroll
                    for (int k = 0; k < UN; k++) {
                        const int i = i0 + k * XE_THREADS;
                        if (i >= nit) continue;
                        if (p.out_i32) {
                            int2 o = v[k];
                            if (p.accumulate) {
                                o.x += oi[k].x;
                                o.y += oi[k].y;
                            }
                            reinterpret_cast<int2 *>(p.out_i32)[gbase + i] = o;
                        }
                        if (p.out_f32) {
                            float2 o = make_float2((float)v[k].x * p.scale, (float)v[k].y * p.scale);
                            if (p.accumulate) {
                                o.x += of[k].x;
                                o.y += of[k].y;
                            }
                            p.out_f32[gbase + i] = o;
                        }
                    }
                }
                if (fix && tid == 0) fx.count[grp_e] = 0;
            }
            tm_worker_sync();
            TM_STAMP(3);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(FC * 32));
    }
}

typedef void (*tm_kernel_t)(XeParams, TmFix, const CUtensorMap);
inline tm_kernel_t tm_kernel(int npol, int fc)
{
    if (npol == 1) return fc == 8 ? &k_xengine_tma<1, 8> : &k_xengine_tma<1, 16>;
    return fc == 8 ? &k_xengine_tma<2, 8> : &k_xengine_tma<2, 16>;
}

// ---- host: tensor map over the caller's [t][station][row bytes] integration buffer ----
typedef CUresult (*tm_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                 const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline tm_encode_fn tm_encoder()
{
    static tm_encode_fn fn = [] {
        void *f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            f = nullptr;
        return (tm_encode_fn)f;
    }();
    return fn;
}
// dims (innermost first): row bytes | time (stride A*rowb) | station (stride rowb); box IB x 32 x ASTN
inline bool tm_make_map(CUtensorMap *tm, const void *base, long rowb, int A, int T, int npol, int fc, int l2promo)
{
    tm_encode_fn enc = tm_encoder();
    if (!enc) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)rowb, (cuuint64_t)T, (cuuint64_t)A};
    const cuuint64_t strides[2] = {(cuuint64_t)rowb * A, (cuuint64_t)rowb};
    const int ib = fc * npol * 2;
    const cuuint32_t box[3] = {(cuuint32_t)ib, (cuuint32_t)(512 / fc), (cuuint32_t)(32 / npol)};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void *>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE,
                     ib == 16 ? CU_TENSOR_MAP_SWIZZLE_NONE : ib == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_64B,
                     (CUtensorMapL2promotion)l2promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

} // namespace
