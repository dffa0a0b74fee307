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

#ifndef CLB_TM_NR
#define CLB_TM_NR 3
#define CLB_TM_NI 3
#define CLB_TM_IMGB (16 * (2048 + 64))
#endif
constexpr int TM_NR = CLB_TM_NR;               // raw ring depth
constexpr int TM_NI = CLB_TM_NI;               // image ring depth
constexpr int TM_RAWB = 32 * 1024;             // bytes per raw box: 32 t x (32 stations x 32 B | 16 x 64 B)
constexpr int TM_IMGB = CLB_TM_IMGB;           // bytes reserved per stage image (FC channels x KT steps x 64 rows + skew)
// the image ring doubles as the epilogue's staging area: 16 channels x (544 + 4) visibilities x 8 B at most
static_assert(TM_NI * TM_IMGB >= 16 * 548 * 8, "epilogue staging does not fit the image ring");
static_assert(TM_NR * TM_RAWB >= 16 * 548 * 8, "cluster exchange does not fit the raw ring");
constexpr int TM_SMEM = 1024 + TM_NR * TM_RAWB + TM_NI * TM_IMGB + 256;
constexpr int TM_THREADS = XE_THREADS + 64;    // 16 transpose/epilogue warps + MMA warp + TMA warp

__device__ __forceinline__ void tm_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(bar)),
                 "r"(bytes)
                 : "memory");
}
// box of the 4-D map (row bytes | time | station | integration of a batch)
__device__ __forceinline__ void tm_load_4d(uint32_t dst, const CUtensorMap *tm, uint64_t *bar, int c0, int c1, int c2, int c3)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(tm), "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
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

__device__ __forceinline__ uint32_t tm_mapa(uint32_t saddr, int rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void tm_mbar_arrive_remote(uint32_t cluster_addr)
{
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tm_mbar_wait_cluster(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "TMW_%=:\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra TMD_%=;\n\t"
        "bra TMW_%=;\n\t"
        "TMD_%=:\n\t}" ::"r"((uint32_t)__cvta_generic_to_shared(bar)),
        "r"(parity)
        : "memory");
}

// Gather write-out, 16 B at a time: pairs of visibilities (npp even, so a pair never straddles two channels and
// both the staging and the output addresses are 16 B aligned).  One multicast store per pair when the matrices are
// bound to an NVSwitch multicast object (the switch replicates it to every rank), else one peer store per rank.
template <int NPOL>
__device__ __forceinline__ void tm_write_out_gather16(const XeParams &p, const int2 *stg, const int2 *recv, int nrecv,
                                                      int rstride, int spp, int npp, int f0, int nch, int tid)
{
    const float scale = p.scale;
    const int hpp = npp >> 1, hsp = spp >> 1;                // pairs per channel in the output / in the staging
    const int npair = nch * hpp, pad = hsp - hpp;
    const int4 *stg4 = reinterpret_cast<const int4 *>(stg), *recv4 = reinterpret_cast<const int4 *>(recv);
    const int rs4 = rstride >> 1;
    const long obase = (p.gather_off + (long)f0 * npp) >> 1;  // in float4 units
    constexpr int UN = 4;
    int r = tid, si = tid;
    while (r >= hpp) {
        r -= hpp;
        si += pad;
    }
    for (int i0 = tid; i0 < npair; i0 += UN * XE_THREADS) {
        int sk[UN];
        int4 v[UN];
#pragma unroll
        for (int k = 0; k < UN; k++) {
            sk[k] = si;
            r += XE_THREADS;
            si += XE_THREADS;
            while (r >= hpp) {
                r -= hpp;
                si += pad;
            }
        }
#pragma unroll
        for (int k = 0; k < UN; k++) v[k] = (i0 + k * XE_THREADS < npair) ? stg4[sk[k]] : make_int4(0, 0, 0, 0);
#pragma unroll 1
        for (int s = 0; s < nrecv; s++) {
            int4 w[UN];
#pragma unroll
            for (int k = 0; k < UN; k++)
                w[k] = (i0 + k * XE_THREADS < npair) ? recv4[s * rs4 + sk[k]] : make_int4(0, 0, 0, 0);
#pragma unroll
            for (int k = 0; k < UN; k++) {
                v[k].x += w[k].x;
                v[k].y += w[k].y;
                v[k].z += w[k].z;
                v[k].w += w[k].w;
            }
        }
        float4 o[UN];
#pragma unroll
        for (int k = 0; k < UN; k++)
            o[k] = make_float4((float)v[k].x * scale, (float)v[k].y * scale, (float)v[k].z * scale, (float)v[k].w * scale);
        if (p.gather_mc != nullptr) {
            float4 *dst = reinterpret_cast<float4 *>(p.gather_mc) + obase;
#pragma unroll
            for (int k = 0; k < UN; k++)
                if (i0 + k * XE_THREADS < npair)
                    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + i0 + k * XE_THREADS),
                                 "f"(o[k].x), "f"(o[k].y), "f"(o[k].z), "f"(o[k].w)
                                 : "memory");
        } else {
            for (int g = 0; g < p.ngather; g++) {
                float4 *dst = reinterpret_cast<float4 *>(p.gather[g]) + obase;
#pragma unroll
                for (int k = 0; k < UN; k++)
                    if (i0 + k * XE_THREADS < npair) dst[i0 + k * XE_THREADS] = o[k];
            }
        }
    }
}

// Final visibilities of `nch` channels starting at channel f0: staged partial sums (`stg`, channel
// stride spp) plus `nrecv` received blocks of the same shape, written in output order
// [channel][baseline][pol^2] as int32 pairs and/or scaled floats (= or +=).
template <int NPOL>
__device__ __forceinline__ void tm_write_out(const XeParams &p, const int2 *stg, const int2 *recv, int nrecv,
                                             int rstride, int spp, int npp, int f0, int nch, int tid, long kofs = 0)
{
    if (p.ngather > 0 && (npp & 1) == 0 && ((p.gather_off + (long)f0 * npp) & 1) == 0 && p.out_i32 == nullptr &&
        p.out_f32 == nullptr) {
        tm_write_out_gather16<NPOL>(p, stg, recv, nrecv, rstride, spp, npp, f0, nch, tid);
        return;
    }
    int2 *const oi = reinterpret_cast<int2 *>(p.out_i32) + kofs + (long)f0 * npp;      // kofs: integration of a batch
    float2 *const of = p.out_f32 + kofs + (long)f0 * npp;
    const bool has_i = p.out_i32 != nullptr, has_f = p.out_f32 != nullptr;
    const bool rmw = p.accumulate != 0, redo = p.split && p.nslice <= 1;
    const float scale = p.scale;
    const int nit = nch * npp, pad = spp - npp;
    // item i = (channel, visibility r): the staging stride per channel is spp, the output's is npp.
    // The epilogue runs once per CTA with only 4 warps per scheduler, so it is bound by the LENGTH of
    // its dependent chains: batches of UN items, every load of a batch issued before the first use.
    constexpr int UN = 5;
    int r = tid, si = tid;
    while (r >= npp) {
        r -= npp;
        si += pad;
    }
    for (int i0 = tid; i0 < nit; i0 += UN * XE_THREADS) {
        int sk[UN];
        int2 v[UN];
#pragma unroll
        for (int k = 0; k < UN; k++) {
            sk[k] = si;
            r += XE_THREADS;
            si += XE_THREADS;
            while (r >= npp) {
                r -= npp;
                si += pad;
            }
        }
#pragma unroll
        for (int k = 0; k < UN; k++) v[k] = (i0 + k * XE_THREADS < nit) ? stg[sk[k]] : make_int2(0, 0);
#pragma unroll 1
        for (int s = 0; s < nrecv; s++) {
            int2 w[UN];
#pragma unroll
            for (int k = 0; k < UN; k++)
                w[k] = (i0 + k * XE_THREADS < nit) ? recv[s * rstride + sk[k]] : make_int2(0, 0);
#pragma unroll
            for (int k = 0; k < UN; k++) {
                v[k].x += w[k].x;
                v[k].y += w[k].y;
            }
        }
        if (redo || rmw) {                               // rare paths
#pragma unroll
            for (int k = 0; k < UN; k++) {
                const int i = i0 + k * XE_THREADS;
                if (i < nit) {
                    if (redo) {                          // stream-K: partial groups meet in memory
                        atomicAdd(&oi[i].x, v[k].x);
                        atomicAdd(&oi[i].y, v[k].y);
                    } else {
                        if (has_i) {
                            const int2 q = oi[i];
                            oi[i] = make_int2(v[k].x + q.x, v[k].y + q.y);
                        }
                        if (has_f) {
                            const float2 q = of[i];
                            of[i] = make_float2((float)v[k].x * scale + q.x, (float)v[k].y * scale + q.y);
                        }
                    }
                }
            }
            continue;
        }
        if (has_i) {
#pragma unroll
            for (int k = 0; k < UN; k++)
                if (i0 + k * XE_THREADS < nit) oi[i0 + k * XE_THREADS] = v[k];
        }
        if (has_f) {
#pragma unroll
            for (int k = 0; k < UN; k++)
                if (i0 + k * XE_THREADS < nit)
                    of[i0 + k * XE_THREADS] = make_float2((float)v[k].x * scale, (float)v[k].y * scale);
        }
        // fused all-gather: the same block goes into every rank's full matrix (peer-mapped over NVLink)
        for (int g = 0; g < p.ngather; g++) {
            float2 *dst = p.gather[g] + p.gather_off + (long)f0 * npp;
#pragma unroll
            for (int k = 0; k < UN; k++)
                if (i0 + k * XE_THREADS < nit)
                    dst[i0 + k * XE_THREADS] = make_float2((float)v[k].x * scale, (float)v[k].y * scale);
        }
    }
}

// packed 4-bit samples (hi nibble re, lo nibble im; CharToComplex LUT lib/clXEngine_impl.cc:833: 4-bit two's complement
// with -8 read as 0): 4 time steps of one sample lane -> the 4 re bytes and the 4 im bytes, SIMD within a register
__device__ __forceinline__ uint32_t tm_nib4(uint32_t n)          // n: four nibbles 0..15, one per byte
{
    const uint32_t z = n ^ 0x08080808u;                               // 0 exactly where the nibble is 8
    const uint32_t r = ((z | 0x80808080u) - 0x08080808u) ^ 0x80808080u;   // per byte (n ^ 8) - 8 = sign extension
    const uint32_t nz = (((z + 0x7F7F7F7Fu) | z) >> 7) & 0x01010101u;     // 1 where z != 0
    return r & (nz * 0xFFu);
}

// FC = channels per CTA: 16 fills the 512 TMEM columns; 8 (256 columns) lets 1024 channels spread over
// 128 CTAs WITHOUT slicing time, i.e. without any cross-CTA reduction.  A raw box is always 32 KiB:
// KT = 512 / FC time steps (KS = KT / 32 MMA k-steps per stage).
// PK: the input is packed 4 bit (one byte per sample): the boxes are half as large and the transpose stage expands the
// nibbles between ldmatrix and stmatrix (one load -> a re and an im store), so packed data crosses HBM once, as 1 B/sample.
template <int NPOL, int FC, bool PK = false>
__global__ void __launch_bounds__(TM_THREADS, 1)
k_xengine_tma(XeParams p, const __grid_constant__ CUtensorMap tmap)
{
    constexpr int IB = FC * NPOL * (PK ? 1 : 2);             // bytes per (t, station) run: 16 | 32 | 64
    constexpr int NH = IB / 16;                              // 16 B chunks per run
    constexpr int KT = 512 / FC;                             // time steps per stage
    constexpr int KS = KT / 32;                              // MMA k-steps per stage
    constexpr int ASTN = 32 / NPOL;                          // stations per box
    constexpr int CSK = KS * 2048 + ((NPOL == 1) ? 32 : 64); // channel image stride: the skew spreads the STSM banks
    constexpr int RAWB = ASTN * KT * IB;                     // bytes per raw box: 32 KiB (16 KiB packed)
    constexpr int UNITS = RAWB / 512;                        // ldmatrix.x2 units per stage
    constexpr int UPW = UNITS / XE_WARPS;                    // ... per warp
    static_assert(FC * CSK <= TM_IMGB, "stage image does not fit");
    static_assert(RAWB <= TM_RAWB && IB >= 16 && UPW >= 1, "raw box shape");

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
    uint64_t *pdone = tfree + 1;                     //       every peer of the cluster is past its main loop
    uint64_t *pdata = tfree + 2;                     //       every peer's partial sums have arrived
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tfree + 3);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int CS = CSK;
    const int nbl = p.A * (p.A + 1) / 2;
    const int npp = nbl * NPOL * NPOL;                       // visibilities per channel
    const int spp = (npp + 5) & ~1;                          // staging stride per channel (int2): breaks the bank tie, even -> 16 B blocks
    const int ngroups = (p.F + FC - 1) / FC;
    const int nst = (p.T + KT - 1) / KT;

    // Time-sliced launches (nslice > 1) are CLUSTER launches: the nslice CTAs of a cluster share one
    // channel group, each integrates a slice of the time steps, and the partial visibilities meet
    // through distributed shared memory (no atomics, no workspace, no second kernel).
    const bool clustered = p.nslice > 1;
    int s0, s1, slice = 0, grp_c = 0, kb = 0;
    if (p.nslice > 0) {
        // batched launches (p.nbatch integrations back to back in memory, one grid): CTA = (integration, group, slice)
        const int gidx = clustered ? blockIdx.x / p.nslice : blockIdx.x;
        kb = gidx / ngroups;
        grp_c = gidx - kb * ngroups;
        slice = clustered ? blockIdx.x % p.nslice : 0;
        const int len = (nst + p.nslice - 1) / p.nslice;
        s0 = grp_c * nst + slice * len;
        s1 = min(grp_c * nst + nst, s0 + len);
    } else if (p.split) {
        const long total = (long)ngroups * nst;
        s0 = (int)(total * blockIdx.x / gridDim.x);
        s1 = (int)(total * (blockIdx.x + 1) / gridDim.x);
    } else {
        // whole groups per CTA, persistent: CTA c takes every gridDim.x-th VIRTUAL group (integration of the batch,
        // channel group); the pipeline below runs across group boundaries, so the write-out of one group overlaps
        // the loads of the next.  Strided, not a contiguous share: the CTAs that run together then work on neighbouring channel groups of the
        // same integration, i.e. on neighbouring 32 B runs of the same (t, station) rows, which is what lets the DRAM
        // controllers serve them from open pages (a contiguous share per CTA scatters the 148 streams over as many
        // rows: measured 43 instead of 12 us per integration).  Local stage index n -> local group n / nst ->
        // virtual group blockIdx.x + (n / nst) * gridDim.x.
        const int vg = ngroups * p.nbatch;
        const int mine = vg > (int)blockIdx.x ? (vg - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
        s0 = 0;
        s1 = mine * nst;
    }
    // a time slice may be empty (more slices than stages); it still takes part in the slice fix-up
    const int nstages = max(0, s1 - s0);
    if (nstages == 0 && !clustered) return;

    // Programmatic dependent launch: the next kernel of the stream may be scheduled as soon as SMs free up;
    // it runs its own prologue and then waits (griddepcontrol.wait below) for this grid to complete.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
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
        tc_mbar_init(pdone, max(p.nslice - 1, 1));
        tc_mbar_init(pdata, max(p.nslice - 1, 1) * XE_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // everything above overlapped the previous grid's tail; from here on global memory is touched
    // ... unless this launch neither reads nor accumulates into anything the previous grid writes (plain `=` result,
    // read-only input): then the CTAs of consecutive integrations simply interleave as SMs free up, and the previous
    // grid's epilogue and tail overlap this grid's first loads (its completion is still ordered behind the previous one's)
    if (!p.pdl_nowait) asm volatile("griddepcontrol.wait;" ::: "memory");
    // peers signal each other's mbarriers: those must be initialised cluster-wide before anyone proceeds
    if (clustered) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == XE_WARPS + 1) {
        // ---- TMA producer ----
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
            int grp = s0 / nst, st = s0 - grp * nst, kl = kb;
            const bool strided = p.nslice <= 0 && !p.split;
            int gv = blockIdx.x;                               // strided: virtual group -> (integration, group)
            if (strided) {
                kl = gv / ngroups;
                grp = gv - kl * ngroups;
            }
            for (int n = 0; n < nstages; n++) {
                const int r = n % TM_NR;
                if (n >= TM_NR) tc_mbar_wait(&rempty[r], (uint32_t)(n / TM_NR - 1) & 1u);
                tm_expect_tx(&rfull[r], RAWB);
                tm_load_4d(raw_addr + r * TM_RAWB, &tmap, &rfull[r], (p.f_off + grp * FC) * NPOL * (PK ? 1 : 2), st * KT, 0, kl);
                if (++st == nst) {
                    st = 0;
                    if (strided) {
                        gv += gridDim.x;
                        kl = gv / ngroups;
                        grp = gv - kl * ngroups;
                    } else {
                        grp++;
                    }
                }
            }
        }
        __syncwarp();
        if (clustered) {
            // the raw ring is idle once the last boxes are transposed: tell the peers they may send
            for (int n = max(0, nstages - TM_NR); n < nstages; n++)
                tc_mbar_wait(&rempty[n % TM_NR], (uint32_t)(n / TM_NR) & 1u);
            if (lane < p.nslice && lane != slice)
                tm_mbar_arrive_remote(tm_mapa((uint32_t)__cvta_generic_to_shared(pdone), lane));
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
            constexpr int UPS = UNITS / ASTN;                    // units per station
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
                // unpacked: byte = (channel, pol, re|im); packed: byte = (channel, pol), its re row (im row = +16 B)
                const int f = PK ? ((NPOL == 1) ? B : (B >> 1)) : ((NPOL == 1) ? (B >> 1) : (B >> 2));
                const int v = PK ? ((NPOL == 1) ? s : 2 * s + (B & 1)) : ((NPOL == 1) ? s : 2 * s + ((B >> 1) & 1));
                const int m = PK ? 2 * v : 2 * v + (B & 1);
                st_off[i] = (uint32_t)(f * CS + (kc >> 1) * 2048 + (m >> 3) * 256 + (kc & 1) * 128 + (m & 7) * 16);
            }
        }
        const int tid = threadIdx.x;                               // 0..511 (worker warps come first)
        for (int n = 0; n < max(nstages, 1); n++) {
            const int sg = s0 + n, b = n % TM_NI, r = n % TM_NR;
            if (nstages > 0) {
                tc_mbar_wait(&rfull[r], (uint32_t)(n / TM_NR) & 1u);
                if (n >= TM_NI) tc_mbar_wait(&empty[b], (uint32_t)(n / TM_NI - 1) & 1u);   // MMAs of stage n-NI are done
                const uint32_t src = raw_addr + r * TM_RAWB, dst = img_addr + b * TM_IMGB;
                {
                    uint32_t v[UPW][4];
#pragma unroll
                    for (int i = 0; i < UPW; i++) tm_ldsm_t8(src + ld_off[i], v[i]);
                    if constexpr (PK) {
#pragma unroll
                        for (int i = 0; i < UPW; i++) {
                            uint32_t re[4], im[4];
#pragma unroll
                            for (int k = 0; k < 4; k++) {
                                re[k] = tm_nib4((v[i][k] >> 4) & 0x0F0F0F0Fu);
                                im[k] = tm_nib4(v[i][k] & 0x0F0F0F0Fu);
                            }
                            tm_stsm(dst + st_off[i], re);
                            tm_stsm(dst + st_off[i] + 16, im);
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < UPW; i++) tm_stsm(dst + st_off[i], v[i]);
                    }
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
            // ---- epilogue.  All MMAs of the group are done and no worker is ahead of this point, so the
            // image ring is free: the 16 channels are combined (re/im rows sit in adjacent lanes), staged
            // there in OUTPUT order [channel][baseline][pol^2](re, im) and leave as ONE contiguous block
            // (the output is channel-major, lib/clXEngine_impl.cc:799-806).
            int grp_e = grp_c, kb_e = kb;
            if (p.nslice <= 0) {
                grp_e = sg / nst;
                if (!p.split) {                                        // strided virtual groups
                    const int gv = blockIdx.x + grp_e * gridDim.x;
                    kb_e = gv / ngroups;
                    grp_e = gv - kb_e * ngroups;
                }
            }
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
                        if (half == 1 && wq < 2) break;                // columns 16..31 only pair with rows >= 16
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
            if (clustered) break;                                      // the exchange below needs every thread
            tm_worker_sync();
            tm_write_out<NPOL>(p, stg, nullptr, 0, 0, spp, npp, f0, min(FC, p.F - f0), tid, (long)kb_e * p.F * npp);
            tm_worker_sync();
        }
    }
    if (clustered && warp < XE_WARPS) {
        // ---- cluster exchange: rank r finalises channels [r*nc, (r+1)*nc) of the group.  A rank's raw ring
        // is idle once its last box is transposed (its TMA lane tells the peers through their `pdone`
        // mbarriers); it then receives the peers' staged partial sums of its channels (16 B remote stores),
        // each sending warp signalling the receiver's `pdata`.  One-way signals instead of cluster-wide
        // barriers (~1 us each here).  Measured slower: one cp.async.bulk shared::cta -> shared::cluster
        // copy per peer (34 KiB took ~1.3 us to land). ----
        const int CL = p.nslice, nc = FC / CL, blk = nc * spp;         // int2 items per (rank, peer) block
        const int tid = threadIdx.x;
        tm_worker_sync();                                              // staging complete
        tm_mbar_wait_cluster(pdone, 0);
        for (int q = 0; q < CL; q++) {
            if (q == slice) continue;
            const int slot = slice < q ? slice : slice - 1;
            const uint32_t remote = tm_mapa(raw_addr + (uint32_t)(slot * blk * 8), q);
            const int4 *src = reinterpret_cast<const int4 *>(stg + q * blk);
            constexpr int UN = 5;
            for (int i0 = tid; i0 < blk / 2; i0 += UN * XE_THREADS) {
                int4 v[UN];
#pragma unroll
                for (int k = 0; k < UN; k++)
                    v[k] = (i0 + k * XE_THREADS < blk / 2 && nstages > 0) ? src[i0 + k * XE_THREADS] : make_int4(0, 0, 0, 0);
#pragma unroll
                for (int k = 0; k < UN; k++)
                    if (i0 + k * XE_THREADS < blk / 2)
                        asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(
                                         remote + (uint32_t)(i0 + k * XE_THREADS) * 16u),
                                     "r"(v[k].x), "r"(v[k].y), "r"(v[k].z), "r"(v[k].w)
                                     : "memory");
            }
            __syncwarp();
            if (lane == 0) tm_mbar_arrive_remote(tm_mapa((uint32_t)__cvta_generic_to_shared(pdata), q));
        }
        const int f0 = grp_c * FC + slice * nc;
        if (nstages == 0) {
            for (int i = tid; i < blk; i += XE_THREADS) stg[slice * blk + i] = make_int2(0, 0);
            tm_worker_sync();
        }
        tm_mbar_wait_cluster(pdata, 0);
        tm_write_out<NPOL>(p, stg + slice * blk, reinterpret_cast<const int2 *>(sptr), CL - 1, blk, spp, npp, f0,
                           min(nc, p.F - f0), tid, (long)kb * p.F * npp);
    }
    if (p.ngather > 0 && p.gather_counter != nullptr && warp < XE_WARPS) {
        // ---- "my slab is complete everywhere": the CTA's workers meet at a barrier (their peer stores are then
        // ordered before it), ONE thread makes them visible system-wide (fence cumulativity) and counts the CTA in;
        // the last CTA releases this rank's flag on every rank (one multicast store, or one peer store per rank).
        // A consumer acquires the flags on the device (k_gather_wait): no host barrier between the kernel and its
        // readers.  (A fence by all 512 threads instead of one cost 15 us per launch.) ----
        tm_worker_sync();
        if (threadIdx.x == 0) {
            if (p.gather_fence_gpu) asm volatile("fence.acq_rel.gpu;" ::: "memory");
            else asm volatile("fence.acq_rel.sys;" ::: "memory");
            const unsigned old = atomicAdd(p.gather_counter, 1u);
            if (old == gridDim.x - 1) {
                atomicExch(p.gather_counter, 0u);                      // ready for the next launch
                asm volatile("fence.acq_rel.sys;" ::: "memory");
                if (p.gather_flag_mc != nullptr) {
                    asm volatile("multimem.st.release.sys.global.u32 [%0], %1;" ::"l"(p.gather_flag_mc), "r"(p.gather_epoch) : "memory");
                } else {
                    for (int g = 0; g < p.ngather; g++)
                        if (p.gather_flag[g] != nullptr)
                            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p.gather_flag[g]), "r"(p.gather_epoch) : "memory");
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(FC * 32));
    }
}

typedef void (*tm_kernel_t)(XeParams, const CUtensorMap);
inline tm_kernel_t tm_kernel(int npol, int fc, bool packed = false)
{
    if (packed) return npol == 1 ? &k_xengine_tma<1, 16, true> : &k_xengine_tma<2, 16, true>;
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
// dims (innermost first): row bytes | time (stride A*rowb) | station (stride rowb) | integration of a batch (stride
// T*A*rowb); box IB x KT x ASTN x 1.  Whatever a box covers beyond a dimension is zero-filled, per dimension.
inline bool tm_make_map(CUtensorMap *tm, const void *base, long rowb, int A, int T, int npol, int fc, int l2promo, int nbatch,
                        bool packed = false)
{
    tm_encode_fn enc = tm_encoder();
    if (!enc) return false;
    const cuuint64_t dims[4] = {(cuuint64_t)rowb, (cuuint64_t)T, (cuuint64_t)A, (cuuint64_t)nbatch};
    const cuuint64_t strides[3] = {(cuuint64_t)rowb * A, (cuuint64_t)rowb, (cuuint64_t)rowb * A * T};
    const int ib = fc * npol * (packed ? 1 : 2);
    const cuuint32_t box[4] = {(cuuint32_t)ib, (cuuint32_t)(512 / fc), (cuuint32_t)(32 / npol), 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<void *>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE,
                     ib == 16 ? CU_TENSOR_MAP_SWIZZLE_NONE : ib == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_64B,
                     (CUtensorMapL2promotion)l2promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

} // namespace
