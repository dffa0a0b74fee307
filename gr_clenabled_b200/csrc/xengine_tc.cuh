// xengine_tc.cuh -- clXEngine on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// Same decomposition as the mma.sync kernel in xengine.cu (per channel G = Z Z^T with
// Z[(input,re|im)][t] int8, exact s32 accumulation) for up to 32 inputs x polarisations:
//   * operands: the transposed int8 stage image is written straight into the UMMA
//     canonical K-major (no-swizzle) shared-memory layout -- core matrix = 8 rows x 16 B,
//     K chunks 128 B apart (LBO), row groups 256 B apart (SBO); rows are ordered
//     m = 2*input + (re|im).  One descriptor serves as A and as B.
//   * MMA: one elected thread issues, per 32-step stage, one
//     tcgen05.mma.cta_group::1.kind::i8 (M=64, N=64, K=32) per channel; completion of a
//     stage's MMAs is tracked with tcgen05.commit -> mbarrier, which is what frees the
//     stage buffer for the transposing threads (double buffered).
//   * accumulators: 64x64 s32 per channel in TMEM.  An M=64 accumulator occupies 16
//     lanes of each 32-lane quadrant, so two channels interleave in the same 64 columns
//     (lane offsets 0 and 16) and the 16 channels of a CTA fill all 512 columns.
//     No accumulator registers: the feed prefetches two stages ahead instead of one.
//   * epilogue: tcgen05.ld (32x32b) hands every thread one row of D; the re/im rows of an
//     input sit in adjacent lanes, one shuffle per column pair forms
//     Re = D[re,re] + D[im,im],  Im = D[im,re] - D[re,im]   (lib/clXEngine_impl.cc:729-736).
// Work split (stream-K over (channel group, stage)), loader, byte transposes and output
// conventions are those of xengine.cu.
#pragma once

namespace {

constexpr int TC_FC = 16;                    // channels per CTA
constexpr int TC_CS = 2048 + 16;             // bytes per channel image; +16 spreads the STS banks
constexpr int TC_ZB = TC_FC * TC_CS;         // bytes per stage buffer
constexpr int TC_SMEM = 2 * TC_ZB + 64;      // + 5 mbarriers + TMEM base slot
constexpr int TC_THREADS = XE_THREADS + 32;  // 16 feed/epilogue warps + 1 MMA warp

// instruction descriptor: D = s32, A = B = signed int8, both K-major, N = 64, M = 64
constexpr uint32_t TC_IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((64u >> 3) << 17) | ((64u >> 4) << 24);

__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t saddr)
{
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) |
           (1ull << 46);                      // version 1 (Blackwell), SWIZZLE_NONE, base offset 0
}
__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(TC_IDESC), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void tc_mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void tc_mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "TCW_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra TCD_%=;\n\t"
        "bra TCW_%=;\n\t"
        "TCD_%=:\n\t}" ::"r"((uint32_t)__cvta_generic_to_shared(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

template <int NPOL>
__global__ void __launch_bounds__(TC_THREADS, 1) k_xengine_tc(XeParams p)
{
    constexpr int FC = TC_FC;
    constexpr int ASTN = 32 / NPOL;                          // stations per stage image (padded)
    constexpr int RUNW = FC * NPOL / 2;                      // 32-bit words per (t, station) run
    constexpr int NQUAD = 8 * ASTN * RUNW;                   // = 2048 four-step quads per stage
    constexpr int QPT = NQUAD / XE_THREADS;                  // = 4
    static_assert(NQUAD % XE_THREADS == 0, "stage must tile the CTA");
    // second sample of a word: the next channel (1 pol) or the Y input of the same channel
    constexpr int ZBD = (NPOL == 1) ? TC_CS : 32;

    extern __shared__ __align__(128) uint8_t tc_smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(tc_smem + 2 * TC_ZB);     // [2]
    uint64_t *empty = full + 2;                                              // [2]
    uint64_t *tfree = full + 4;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(full + 5);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nbl = p.A * (p.A + 1) / 2;
    const long rowb = (long)p.Fstride * NPOL * 2;
    const long frameb = rowb * p.A;
    const int ngroups = (p.F + FC - 1) / FC;
    const int nst = (p.T + XE_TT - 1) / XE_TT;

    int s0, s1;
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

    // ---- one-time setup: TMEM (all 512 columns), mbarriers ----
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(
            (uint32_t)__cvta_generic_to_shared(tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (threadIdx.x == 0) {
        tc_mbar_init(&full[0], XE_WARPS);
        tc_mbar_init(&full[1], XE_WARPS);
        tc_mbar_init(&empty[0], 1);
        tc_mbar_init(&empty[1], 1);
        tc_mbar_init(tfree, XE_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // ---- stage-invariant description of this thread's quads (see xengine.cu) ----
    unsigned qsrc[QPT];
    int qz[QPT], qt[QPT], qwi[QPT];
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
        // byte offset of (channel, row 2v, k word q) in a stage buffer
        qz[i] = ch_a * TC_CS + (v_a >> 2) * 256 + (q >> 2) * 128 + (2 * (v_a & 3)) * 16 + (q & 3) * 4;
        qt[i] = 4 * q;
        qwi[i] = wi;
        qok[i] = s < p.A;
    }

    // Stages are loaded strictly in sequence, so a cursor (group, stage, base pointer)
    // replaces per-stage divisions and 64-bit address rebuilds.
    int ld_grp = s0 / nst, ld_st = s0 - ld_grp * nst;
    auto load_stage = [&](int, uint32_t (&pre)[QPT][4]) {
        const int f0 = ld_grp * FC, st = ld_st;
        const int8_t *sbase = p.in + ((long)(p.f_off + f0)) * NPOL * 2 + (long)st * XE_TT * frameb;
        const int trem = p.T - st * XE_TT;
        const bool full = p.aligned && (f0 + FC <= p.F) && (p.A == ASTN);
        if (++ld_st == nst) {
            ld_st = 0;
            ld_grp++;
        }
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
    // 4x4 byte transposes, then four 32-bit stores per quad.  For one polarisation the
    // lanes of a store instruction differ in (channel pair wi, k word): with the 16 B
    // channel skew, lanes wi >= 4 store their odd channel first so that the 8 channels
    // of an instruction land on 8 different 16 B bank groups.
    auto store_stage = [&](uint8_t *z, const uint32_t (&pre)[QPT][4]) {
#pragma unroll
        for (int i = 0; i < QPT; i++) {
            const uint32_t lo01 = __byte_perm(pre[i][0], pre[i][1], 0x5140);
            const uint32_t hi01 = __byte_perm(pre[i][0], pre[i][1], 0x7362);
            const uint32_t lo23 = __byte_perm(pre[i][2], pre[i][3], 0x5140);
            const uint32_t hi23 = __byte_perm(pre[i][2], pre[i][3], 0x7362);
            const uint32_t a_re = __byte_perm(lo01, lo23, 0x5410), a_im = __byte_perm(lo01, lo23, 0x7632);
            const uint32_t b_re = __byte_perm(hi01, hi23, 0x5410), b_im = __byte_perm(hi01, hi23, 0x7632);
            const bool swap = (NPOL == 1) && (qwi[i] & 4);
            uint8_t *d0 = z + qz[i] + (swap ? ZBD : 0);
            uint8_t *d1 = z + qz[i] + (swap ? 0 : ZBD);
            *reinterpret_cast<uint32_t *>(d0) = swap ? b_re : a_re;
            *reinterpret_cast<uint32_t *>(d0 + 16) = swap ? b_im : a_im;
            *reinterpret_cast<uint32_t *>(d1) = swap ? a_re : b_re;
            *reinterpret_cast<uint32_t *>(d1 + 16) = swap ? a_im : b_im;
        }
    };

    // ---- warp-specialised pipeline -------------------------------------------------
    //   warps 0..15 : feed (global -> registers -> transposed K-major image) + epilogue
    //   warp 16     : one lane issues the UMMAs
    //   full[b]  (16 arrivals)  : stage image b is complete and visible to the tensor core
    //   empty[b] (tcgen05.commit): the MMAs that read image b -- and all earlier ones -- are done
    //   tfree    (16 arrivals)  : the epilogue has drained TMEM, a new group may overwrite it
    const int nstages = s1 - s0;
    const uint32_t zaddr = (uint32_t)__cvta_generic_to_shared(tc_smem);

    if (warp == XE_WARPS) {
        if (lane == 0) {
            int groups_done = 0;
            for (int n = 0; n < nstages; n++) {
                const int sg = s0 + n, b = n & 1;
                const bool group_first = (n == 0) || (sg % nst == 0);
                tc_mbar_wait(&full[b], (uint32_t)(n >> 1) & 1u);
                if (group_first && n > 0) {
                    tc_mbar_wait(tfree, (uint32_t)(groups_done - 1) & 1u);
                }
                tc_fence_after();
#pragma unroll
                for (int ch = 0; ch < FC; ch++) {
                    const uint64_t d = tc_smem_desc(zaddr + b * TC_ZB + ch * TC_CS);
                    tc_mma_i8(tmem_base + (((uint32_t)(ch & 1) * 16u) << 16) + (uint32_t)(ch >> 1) * 64u, d, d,
                              group_first ? 0u : 1u);
                }
                tc_commit(&empty[b]);
                if ((n + 1 == nstages) || ((sg + 1) % nst == 0)) groups_done++;
            }
        }
    } else {
        uint32_t pre0[QPT][4], pre1[QPT][4];        // stages n even / n odd, two stages ahead
        load_stage(s0, pre0);
        if (nstages > 1) load_stage(s0 + 1, pre1);
        // Whole-row L2 prefetch (time-sliced mode): the CTAs of a slice walk the same
        // (t, station) rows; each one asks L2 for a share of the COMPLETE rows (all
        // channels) of the stage PF_AHEAD steps ahead, so DRAM sees row-sized bursts and the
        // 32 B demand loads of every CTA find their sectors in L2.
        constexpr int PF_AHEAD = 3;
        const bool do_pf = p.l2_rows && p.nslice > 0 && (rowb % 16 == 0);
        const int rows_per_stage = XE_TT * p.A;
        const int pf_share = (rows_per_stage + ngroups - 1) / ngroups;
        for (int n = 0; n < nstages; n++) {
            const int sg = s0 + n, b = n & 1;
            if (do_pf && n + PF_AHEAD < nstages && (int)threadIdx.x < pf_share) {
                const int grp = sg / nst, st = sg - grp * nst + PF_AHEAD;
                const int r = grp * pf_share + threadIdx.x;             // row of that stage: (t, station)
                const int t = st * XE_TT + r / p.A;
                if (r < rows_per_stage && t < p.T) {
                    const int8_t *row = p.in + ((long)t * p.A + (r % p.A)) * rowb;
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(row), "r"((uint32_t)rowb) : "memory");
                }
            }
            if (n >= 2) tc_mbar_wait(&empty[b], (uint32_t)((n >> 1) - 1) & 1u);   // MMAs of stage n-2 are done
            if (b == 0) {
                store_stage(tc_smem, pre0);
                if (n + 2 < nstages) load_stage(sg + 2, pre0);
            } else {
                store_stage(tc_smem + TC_ZB, pre1);
                if (n + 2 < nstages) load_stage(sg + 2, pre1);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) tc_mbar_arrive(&full[b]);

            const bool group_done = (n + 1 == nstages) || ((sg + 1) % nst == 0);
            if (!group_done) continue;
            tc_mbar_wait(&empty[b], (uint32_t)(n >> 1) & 1u);       // every MMA of the group has landed
            tc_fence_after();
            const int f0 = (sg / nst) * FC;
            const int wq = warp & 3;
            const int v1 = 8 * wq + ((lane & 15) >> 1), c1 = lane & 1;
            const int st1 = (NPOL == 1) ? v1 : (v1 >> 1);
#pragma unroll 1
            for (int jj = 0; jj < 2; jj++) {
                const int j = (warp >> 2) + 4 * jj;                   // channel pair of this pass
                const int f = f0 + 2 * j + (lane >> 4);
                const bool row_ok = f < p.F && st1 < p.A;
                long ob;                                               // output index for column input 0
                if (NPOL == 1) ob = (long)f * nbl + (long)st1 * (st1 + 1) / 2;
                else ob = 4 * ((long)f * nbl + (long)st1 * (st1 + 1) / 2) + 2 * (v1 & 1);
#pragma unroll 1
                for (int half = 0; half < 2; half++) {
                    uint32_t r[32];
                    tc_ld32(tmem_base + (((uint32_t)wq * 32u) << 16) + (uint32_t)(j * 64 + half * 32), r);
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        const int v2 = 16 * half + i;
                        const int mine = (int)r[2 * i];
                        const int other = __shfl_xor_sync(0xffffffffu, (int)r[2 * i + 1], 1);
                        const int val = c1 ? mine - other : mine + other;
                        const int st2 = (NPOL == 1) ? v2 : (v2 >> 1);
                        const long o = ob + ((NPOL == 1) ? v2 : (4 * st2 + (v2 & 1)));
                        if (row_ok && st2 <= st1) {
                            if (p.split) {
                                atomicAdd(p.out_i32 + 2 * o + c1, val);
                            } else {
                                if (p.out_i32) {
                                    int *d = p.out_i32 + 2 * o + c1;
                                    *d = p.accumulate ? *d + val : val;
                                }
                                if (p.out_f32) {
                                    float *d = reinterpret_cast<float *>(p.out_f32) + 2 * o + c1;
                                    const float v = (float)val * p.scale;
                                    *d = p.accumulate ? *d + v : v;
                                }
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) tc_mbar_arrive(tfree);
        }
    }
    tc_fence_before();
    __syncthreads();

    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
    }
}

} // namespace
