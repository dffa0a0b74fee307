/*
 * clenabled_b200.h -- C ABI of libclenabled_b200.so
 *
 * The drop-in boundary of the B200-native gr-clenabled hot path.  Everything
 * above this header (the gr::clenabled::* block classes in
 * gr_clenabled_b200/host/, the ctypes binding in gr_clenabled_b200/capi.py, a
 * maintainer's own GNU Radio build) talks to the CUDA kernels only through
 * these entry points: plain pointers and sizes, opaque handles, int status.
 * No torch / GNU Radio / C++ types cross it.
 *
 * Every entry point cites the reference interface it replaces
 * (paths relative to the gr-clenabled tree).
 *
 * Conventions
 *   - return 0 on success, <0 on error (CLB200_E*); the message of the last
 *     error raised on the calling thread is clb200_last_error().
 *   - "work" calls take HOST pointers (pageable or pinned), copy through the
 *     handle's pinned rings on the handle's own CUDA streams, and are complete
 *     (outputs written) when they return -- the contract of gr::block::work().
 *   - "launch_device" calls take DEVICE pointers and a cudaStream_t (as
 *     void*), enqueue the same kernels and return without synchronising.  They
 *     are what a device-resident pipeline (and the roofline bench) uses.
 *   - a handle is not re-entrant for work(); setters are mutex-guarded and may
 *     be called from another thread (reference: d_mutex / d_setlock).
 *   - gr_complex == interleaved {float re, im} (include/clenabled/clSComplex.h:12-17).
 */
#ifndef CLENABLED_B200_H
#define CLENABLED_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define CLB200_API __attribute__((visibility("default")))
#else
#define CLB200_API
#endif

/* status codes */
#define CLB200_OK        0
#define CLB200_EINVAL   -1   /* bad argument (the reference throws in the ctor) */
#define CLB200_ECUDA    -2   /* CUDA runtime error (the reference prints + exit(0), GRCLBase.cpp:239-257) */
#define CLB200_ENOMEM   -3
#define CLB200_ESTATE   -4   /* call not valid in the handle's current state */

/* data types: include/clenabled/GRCLBase.h:57-62 (the numeric values of the reference: saved flowgraphs
 * and reference-API callers pass them as plain ints) */
#define CLB200_DTYPE_COMPLEX   1
#define CLB200_DTYPE_FLOAT     2
#define CLB200_DTYPE_INT       3
#define CLB200_DTYPE_SHORT     4   /* reserved: no hot-path block takes int16 */
#define CLB200_DTYPE_BYTE      5   /* clXEngine IChar: interleaved int8 re,im */
#define CLB200_DTYPE_PACKEDXY  6   /* clXEngine packed 4-bit */

/* operator codes: include/clenabled/clMathOpTypes.h:11-20 */
#define CLB200_OP_MULTIPLY       1
#define CLB200_OP_ADD            2
#define CLB200_OP_SUBTRACT       3
#define CLB200_OP_COMPLEX_CONJ   4
#define CLB200_OP_MULTIPLY_CONJ  5
#define CLB200_OP_EMPTY        255
#define CLB200_OP_EMPTY_W_COPY 254

/* FFT direction: clFFT's CLFFT_FORWARD=-1 / CLFFT_BACKWARD=+1 (grc/clenabled_clFFT.block.yml:37-41) */
#define CLB200_FFT_FORWARD  (-1)
#define CLB200_FFT_BACKWARD (1)

/* secondary element-wise kernels */
#define CLB200_UNARY_LOG10            1  /* clLog           lib/clLog_impl.cc:139-148 */
#define CLB200_UNARY_COMPLEX_TO_MAG   2  /* clComplexToMag  lib/clComplexToMag_impl.cc:140-148 */
#define CLB200_UNARY_COMPLEX_TO_ARG   3  /* clComplexToArg  lib/clComplexToArg_impl.cc:139-151 */

typedef struct clb200_block *clb200_handle;

/* ---------------------------------------------------------------- runtime -- */
/* replaces GRCLBase::InitOpenCL device discovery (lib/GRCLBase.cpp:17-369)     */
CLB200_API const char *clb200_version(void);
CLB200_API const char *clb200_last_error(void);
CLB200_API int clb200_device_count(void);                 /* <0 on error, 0 = no GPU */
CLB200_API int clb200_device_name(int device, char *buf, int buflen);
CLB200_API int clb200_device_sm_count(int device);
/* (openCLPlatformType, devSelector, platformId, devId) -> CUDA ordinal
 * (GRCLBase.h:64-70, GRCLBase.cpp:115-188): devId when devSelector==2 else 0 */
CLB200_API int clb200_select_device(int platform_type, int dev_selector, int platform_id, int dev_id);
CLB200_API int clb200_destroy(clb200_handle h);           /* any handle; GRCLBase::cleanup/stop */
/* counters since creation: bytes H2D, bytes D2H, kernel launches */
CLB200_API int clb200_get_counters(clb200_handle h, uint64_t *h2d, uint64_t *d2h, uint64_t *launches);

/* setDebug of the reference factories (GR_LOG_INFO of the device and kernel, e.g. lib/clXEngine_impl.cc:700-703):
 * describe() gives the kernel variant and launch geometry the handle was configured with; set_debug(1)
 * makes the host path print one line per work() call (items, chunks, microseconds) on stderr.           */
CLB200_API int clb200_describe(clb200_handle h, char *buf, int buflen);
CLB200_API int clb200_set_debug(clb200_handle h, int on);

/* Measured FP32 pipe peaks of the device (a ~10 ms probe): FFMA in TFLOP/s (2 flop per lane and clock), plain and
 * packed (add.f32x2) additions in TFLOP/s (1 flop per lane and clock).  The compute-side roofline denominators of the
 * FP32-bound kernels (FFT filter, time-domain FIR) in bench.py.                                         */
CLB200_API int clb200_probe_fp32(int device, double *ffma_tflops, double *fadd_tflops, double *fadd2_tflops);

/* Optional: page-lock a caller-owned host range (e.g. a GNU Radio circular buffer,
 * once, in start()) so that work() calls on it skip the staging copy and are DMA'd /
 * read by the kernels in place.  The range must stay valid until unregistered.   */
CLB200_API int clb200_register_host_buffer(void *ptr, size_t bytes);
CLB200_API int clb200_unregister_host_buffer(void *ptr);

/* ------------------------------------------------------------ clMathConst -- */
/* clMathConst::make(idataType,..,fValue,operatorType,..) include/clenabled/clMathConst.h:51
 * work: clMathConst_impl::processOpenCL lib/clMathConst_impl.cc:311-361            */
CLB200_API int clb200_mathconst_create(int dtype, int device, float k, int op, clb200_handle *out);
CLB200_API int clb200_mathconst_set_k(clb200_handle h, float k);    /* clMathConst.h:54 */
CLB200_API float clb200_mathconst_k(clb200_handle h);               /* clMathConst.h:53 */
CLB200_API int clb200_mathconst_work(clb200_handle h, const void *in, void *out, long nitems);
CLB200_API int clb200_mathconst_launch_device(clb200_handle h, const void *d_in, void *d_out,
                                              long nitems, void *stream);

/* --------------------------------------------------------------- clMathOp -- */
/* clMathOp::make(idataType,..,operatorType,..) include/clenabled/clMathOp.h:42
 * work: clMathOp_impl::processOpenCL lib/clMathOp_impl.cc:361-442                  */
CLB200_API int clb200_mathop_create(int dtype, int device, int op, clb200_handle *out);
CLB200_API int clb200_mathop_work(clb200_handle h, const void *a, const void *b, void *c, long nitems);
CLB200_API int clb200_mathop_launch_device(clb200_handle h, const void *d_a, const void *d_b,
                                           void *d_c, long nitems, void *stream);

/* --------------------------------------- secondary element-wise (row M5) -- */
/* clLog::make(..,nValue,kValue) / clComplexToMag / clComplexToArg: 1 in -> 1 out */
CLB200_API int clb200_unary_create(int kind, int device, float n_value, float k_value,
                                   clb200_handle *out);
CLB200_API int clb200_unary_work(clb200_handle h, const void *in, void *out, long nitems);
CLB200_API int clb200_unary_launch_device(clb200_handle h, const void *d_in, void *d_out,
                                          long nitems, void *stream);
/* clSNR (lib/clSNR_impl.cc:105-113): c = fabs(n*log10(a/b)+k), 2 float in -> 1 float out */
CLB200_API int clb200_snr_create(int device, float n_value, float k_value, clb200_handle *out);
CLB200_API int clb200_snr_work(clb200_handle h, const float *a, const float *b, float *c, long nitems);
/* clComplexToMagPhase (lib/clComplexToMagPhase_impl.cc:151-165): 1 complex in -> mag, phase */
CLB200_API int clb200_c2magphase_create(int device, clb200_handle *out);
CLB200_API int clb200_c2magphase_work(clb200_handle h, const void *in, float *mag, float *phase,
                                      long nitems);
/* clMagPhaseToComplex (lib/clMagPhaseToComplex_impl.cc:169-192): mag, phase -> complex */
CLB200_API int clb200_magphase2c_create(int device, clb200_handle *out);
CLB200_API int clb200_magphase2c_work(clb200_handle h, const float *mag, const float *phase,
                                      void *out, long nitems);

/* ------------------------------------------------------------------ clFFT -- */
/* clFFT::make(fftSize, clFFTDir, window, idataType, .., num_streams, shift)
 *   (argument order of lib/clFFT_impl.cc:35-36); ctor lib/clFFT_impl.cc:65-151:
 *   window_len must be 0 or fft_size (:74-76 throws).
 * dtype COMPLEX: c32[fft_size] -> c32[fft_size] per item; dtype FLOAT (forward
 * only): f32[fft_size] -> full Hermitian c32[fft_size] spectrum (:556-565,608-630).
 * Unnormalised in both directions (:121-122).  shift: forward swaps the output
 * halves (:594-607), backward swaps the input halves (:548-553); complex data only --
 * with dtype FLOAT the flag is ignored like the reference does (:594).
 * fft_size: powers of two 2 .. 4194304 (up to 16384 one in-SM kernel; above that a four-step
 * decomposition: two passes of those kernels around three transposes); any other length
 * 2 .. 2097152 (clFFT plans take 2^a 3^b 5^c 7^d, :97-100) runs as a chirp-z (Bluestein)
 * transform over the power-of-two kernels.  With an odd length the half swaps leave the
 * last element in place (vlen_2 = fft_size / 2, :81).                                 */
CLB200_API int clb200_fft_create(int fft_size, int dir, const float *window, int window_len,
                                 int dtype, int device, int shift, clb200_handle *out);
/* processOpenCL (lib/clFFT_impl.cc:526-634): nvec items (vectors) of one stream   */
CLB200_API int clb200_fft_work(clb200_handle h, const void *in, void *out, long nvec);
/* the num_streams loop (:537): stream s reads in[s], writes out[s]              */
CLB200_API int clb200_fft_work_streams(clb200_handle h, const void *const *in, void *const *out,
                                       int nstreams, long nvec);
CLB200_API int clb200_fft_launch_device(clb200_handle h, const void *d_in, void *d_out, long nvec,
                                        void *stream);

/* --------------------------------------------------------------- clFilter -- */
/* clFilter::make(..,decimation,taps,nthreads,setDebug,use_time) include/clenabled/clFilter.h:52-53
 * use_time=1: td_FIR_complex (lib/clFilter_impl.cc:162-194, :505-589)
 * use_time=0: FFT filter, fft_filter_ccf sizes (lib/fft_filter.cc:72-97), lib/clFilter_impl.cc:592-681
 * Streaming semantics: y[n] = sum_k taps[k] x[n-k] over the whole stream fed so
 * far (zero initial state); the last ntaps-1 inputs and the decimation phase are
 * kept device-resident between calls, so any chunking gives the same stream.    */
CLB200_API int clb200_filter_create(int device, int decimation, const float *taps, int ntaps,
                                    int use_time, clb200_handle *out);
/* set_taps2 (clFilter.h:55): takes effect on the next work(); history is reset
 * (lib/clFilter_impl.cc:774-789)                                                */
CLB200_API int clb200_filter_set_taps(clb200_handle h, const float *taps, int ntaps);
CLB200_API int clb200_filter_ntaps(clb200_handle h);
CLB200_API int clb200_filter_get_taps(clb200_handle h, float *taps, int cap);   /* clFilter.h:56 */
CLB200_API int clb200_filter_reset(clb200_handle h);
/* fft_filter_ccf::compute_sizes result the reference would use (fft_filter.cc:77-78) */
CLB200_API int clb200_filter_ref_sizes(int ntaps, int *fftsize, int *nsamples);
/* consumes n_in samples, writes *n_out = number of decimated outputs produced    */
CLB200_API int clb200_filter_work(clb200_handle h, const void *in, long n_in, void *out,
                                  long *n_out);
/* device pointers; d_in holds n_in NEW samples (history is inside the handle)   */
CLB200_API int clb200_filter_launch_device(clb200_handle h, const void *d_in, long n_in,
                                           void *d_out, long *n_out, void *stream);

/* ------------------------------------------------- clPolyphaseChannelizer -- */
/* clPolyphaseChannelizer::make(..,taps,buf_items,num_channels,ninputs_per_iter,ch_map,..)
 *   include/clenabled/clPolyphaseChannelizer.h:48-49; ctor checks lib/..._impl.cc:59-62
 * general_work lib/clPolyphaseChannelizer_impl.cc:83-109, kernels :156-177, plan :208-225.
 * in: GNU Radio history layout -- in[0] is ntaps-1 samples in the past; the call
 * reads (niter-1)*R + ntaps samples and writes niter*nmap outputs.              */
CLB200_API int clb200_pfb_create(int device, const float *taps, int ntaps, int buf_items,
                                 int num_channels, int ninputs_per_iter, const int *ch_map,
                                 int nmap, clb200_handle *out);
CLB200_API int clb200_pfb_work(clb200_handle h, const void *in, void *out, long niter);
CLB200_API int clb200_pfb_launch_device(clb200_handle h, const void *d_in, void *d_out, long niter,
                                        void *stream);

/* -------------------------------------------------------------- clXEngine -- */
/* clXEngine::make(.. data_type, polarization, num_inputs, output_format, first_channel,
 *   num_channels, integration, ..) include/clenabled/clXEngine.h:48-52.
 * Input: one integration in the reference host-buffer layout
 *   [t][station][chan][pol] (lib/clXEngine_impl.cc:987-1058); IChar = int8 (re,im),
 *   COMPLEX = c32, PACKEDXY = one byte per sample (hi nibble re, lo nibble im).
 * Output: c32[num_channels][num_baselines][npol*npol], baseline k = s1(s1+1)/2+s2,
 *   s1>=s2, V = sum_t x[s1] conj(x[s2]) (lib/clXEngine_impl.cc:739-810, :204-211).
 * IChar/PACKEDXY accumulate exactly in int32 on the tensor cores; the float
 * result is int32 * (1/127)^2 (resp. (1/7)^2) (CharToComplex :833-866).
 * num_inputs >= 2 else CLB200_EINVAL (std::out_of_range at :106-109).            */
CLB200_API int clb200_xengine_create(int device, int data_type, int npol, int num_inputs,
                                     int num_channels, int integration, clb200_handle *out);
CLB200_API long clb200_xengine_input_bytes(clb200_handle h);    /* one integration  */
CLB200_API long clb200_xengine_output_items(clb200_handle h);   /* matrix_flat_length */
/* xcorrelate(): H2D -> correlate -> D2H (lib/clXEngine_impl.h:150-201, .cc:1234-1299).
 * accumulate!=0 adds into the previous result (pipeline_integration, :785-808).   */
CLB200_API int clb200_xengine_work(clb200_handle h, const void *in, void *out_c32, int accumulate);
/* same, but returns the exact integer accumulators int32[...][2] (IChar/PACKEDXY) */
CLB200_API int clb200_xengine_work_i32(clb200_handle h, const void *in, int32_t *out_i32);
CLB200_API int clb200_xengine_launch_device(clb200_handle h, const void *d_in, void *d_out_c32,
                                            int accumulate, void *stream);
CLB200_API int clb200_xengine_launch_device_i32(clb200_handle h, const void *d_in,
                                                int32_t *d_out_i32, void *stream);
/* Streaming ingest -- the shape of the reference block: work_processor (lib/clXEngine_impl.cc:918-1142) marshals
 * each general_work() call's port vectors into a pinned integration buffer, and a worker (runThread
 * :1234-1299) uploads and correlates a full buffer while the next one fills.  Here:
 *   stream_begin   allocates two pinned + two device integration buffers, a ring of result matrices and
 *                  the copy / compute / read-back streams.  pipeline_integration > 1 sums that many
 *                  integrations on the device before a result is emitted (:785-808, :1249-1284).
 *   push_timesteps `ntime` time steps: ports[s] -> ntime items of station s (item = the caller's channels,
 *                  this handle's slab of them after set_shard); unpacked two-polarisation data comes as
 *                  2*num_inputs ports (X = ports[s], Y = ports[s + num_inputs]) and is interleaved per channel
 *                  (:1010-1057).  Every push is uploaded AT ONCE on the copy stream; a completed integration
 *                  is correlated and its matrix read back asynchronously.  The call only waits for the GPU
 *                  when it is a whole integration behind.  The ports may be reused on return.  Page-locked
 *                  ports (clb200_register_host_buffer) of >= 1 MiB per push are DMA'd in place.
 *   poll_result    copies the oldest finished matrix into out_c32 (*ready = 1) or reports *ready = 0;
 *                  wait != 0 blocks until it is there.  At most result_slots matrices may be pending.
 *   stream_end     drains and frees the stream state (also done by clb200_destroy).                        */
CLB200_API int clb200_xengine_stream_begin(clb200_handle h, int pipeline_integration, int result_slots);
CLB200_API int clb200_xengine_push_timesteps(clb200_handle h, const void *const *ports, int nports, long ntime);
CLB200_API int clb200_xengine_poll_result(clb200_handle h, void *out_c32, int wait, int *ready);
CLB200_API int clb200_xengine_stream_state(clb200_handle h, long *tracker, long *integrations, long *results_pending,
                                           uint64_t *pushes, uint64_t *pushes_blocked);
/* page-locked ports only: the caller promises that the memory handed to push_timesteps stays valid and unchanged until
 * the result of the integration it belongs to has been polled (a capture ring it owns, not a scheduler buffer).  push
 * then returns without waiting for the DMA to have read the ports, so uploads overlap the caller's next push.    */
CLB200_API int clb200_xengine_stream_ports_stable(clb200_handle h, int stable);
CLB200_API int clb200_xengine_stream_end(clb200_handle h);
/* nbatch integrations that lie back to back in device memory (IChar), matrices back to back in d_out_c32: ONE grid
 * of nbatch x (channel groups x time slices) CTAs, so the per-launch fixed cost (first-load latency, exchange,
 * write-out, drain) of one integration overlaps the streaming phase of the others and all SMs stay busy
 * (128 CTAs per 32 x 1024 x 1024 integration leave 20 of 148 SMs idle when launched one by one).       */
CLB200_API int clb200_xengine_launch_device_batch(clb200_handle h, const void *d_in, void *d_out_c32, int nbatch,
                                                  void *stream);
/* channel-sharded variant for multi-GPU: this handle owns channels
 * [chan_first, chan_first+chan_count) of an integration whose host layout has
 * total_channels per station; work() gathers only that slab (cudaMemcpy2D).      */
CLB200_API int clb200_xengine_set_shard(clb200_handle h, int total_channels, int chan_first);
/* Fused all-gather over peer memory (NVLink / NVSwitch), one process per GPU: every rank registers the
 * full visibility matrix of EVERY rank (its own + the peers', opened with clb200_ipc_open); the
 * correlation kernel's epilogue then writes this rank's channel slab straight into all of them
 * (c32[total_channels][baselines][pol^2] at channel offset chan_first), so no collective follows the
 * kernel.  The matrices are complete once every rank's stream has drained (a barrier of the caller's), or -- without
 * any host barrier -- once clb200_xengine_gather_wait has run (set_gather_sync below).
 * Needs clb200_xengine_set_shard first and 16 B aligned input rows (the TMA kernel).               */
#define CLB200_XENGINE_MAX_GATHER 8
CLB200_API int clb200_xengine_set_gather(clb200_handle h, int nranks, void *const *full_out_c32);
/* d_in: this rank's channel slab only, [t][station][shard channels][pol]                            */
CLB200_API int clb200_xengine_launch_device_gather(clb200_handle h, const void *d_in, void *stream);
/* Device-side completion of the fused gather (replaces the caller's barrier) and NVSwitch multicast:
 * flag_arrays[r] = rank r's flag array (uint32[nranks * CLB200_XENGINE_FLAG_STRIDE], zero-initialised, peer-mapped
 * like the matrices).  After its last CTA has stored, a launch release-stores its epoch (1, 2, ... per launch) into
 * word [my_rank * STRIDE] of EVERY rank's array -- by default from the one-warp kernel that clb200_xengine_gather_wait
 * enqueues behind the launch (so the hot kernel carries no system-scope fence), which then acquires all nranks words
 * of the local array: whatever follows it on that stream reads a complete matrix.  EVERY rank calls gather_wait after
 * each gather launch (it is also what announces the rank's own slab); a rank that never does makes the others' wait
 * kernels fail after 10 s instead of hanging.
 * multicast_out / multicast_flags (optional, both or neither): multicast addresses of the matrix and of the flag
 * array bound on every rank (cuMulticast* / torch symmetric memory); the slab and the flag then leave as ONE
 * `multimem.st` each and the switch replicates them, instead of one peer store per rank.               */
#define CLB200_XENGINE_FLAG_STRIDE 32
CLB200_API int clb200_xengine_set_gather_sync(clb200_handle h, int my_rank, void *const *flag_arrays,
                                              void *multicast_out, void *multicast_flags);
CLB200_API int clb200_xengine_gather_wait(clb200_handle h, void *stream);

/* ------------------------------------------------- device memory for peers -- */
/* cudaMalloc'ed buffers whose interprocess handle (64 opaque bytes) other ranks of the same box can
 * open; peer access is enabled on open.  Only the gather above needs them.                          */
CLB200_API int clb200_mem_alloc(int device, size_t bytes, void **dptr);
CLB200_API int clb200_mem_free(int device, void *dptr);
CLB200_API int clb200_mem_copy_to_host(int device, const void *dptr, void *host, size_t bytes);
CLB200_API int clb200_ipc_export(int device, void *dptr, void *handle64);
CLB200_API int clb200_ipc_open(int device, const void *handle64, void **dptr);
CLB200_API int clb200_ipc_close(int device, void *dptr);

/* ================================================================================
 * SURVEY 8(f) "next" rows: the blocks either side of the hot path.
 * ================================================================================ */

/* ---------------------------------------------------------- clXCorrelate -- */
/* clXCorrelate::make(.., num_inputs, signal_length, data_type, data_size, max_search_index,
 *   decim_frames, async) include/clenabled/clXCorrelate.h:56-57.  Input 0 is the reference.
 * Per frame of signal_length samples: magnitude of complex inputs (ComplexToMag,
 * lib/clXCorrelate_impl.cc:915-929), squares (F32Squared :976-980), one normalised correlation
 * factor per shift in [-max_shift, max_shift) (XCorrelate :851-900; -2.0 where the overlap has
 * no energy), then find_max (:1016-1043 + host pass :1371-1413) -> (correlation, corrective lag).
 * max_shift: max_search_index, or 0.7*signal_length made even, rounded up to a power of two
 * (:727-747).  signal_length and max_search_index must be even (:715-724).
 * decim_frames / async are scheduling policies of the block layer above this ABI.          */
CLB200_API int clb200_xcorrelate_create(int device, int num_inputs, int signal_length, int data_type,
                                        int max_search_index, clb200_handle *out);
CLB200_API int clb200_xcorrelate_max_shift(clb200_handle h);
/* in: num_inputs host pointers to signal_length items each; corr/lag: num_inputs-1 entries
 * (the "corrvect" / "corrective_lags" vectors of the PDU, :1585-1593)                         */
CLB200_API int clb200_xcorrelate_work(clb200_handle h, const void *const *in, float *corr, int32_t *lag);
/* d_in: [num_inputs][signal_length] contiguous device buffer; d_corr/d_lag device outputs   */
CLB200_API int clb200_xcorrelate_launch_device(clb200_handle h, const void *d_in, float *d_corr,
                                               int32_t *d_lag, void *stream);
/* the correlation_factors buffer of the last call for non-reference input `signal` (1-based),
 * 2*max_shift floats (what the reference's non-kernel find_max path reads back, :1417-1420)  */
CLB200_API int clb200_xcorrelate_factors(clb200_handle h, int signal, float *out, int cap);

/* ---------------------------------------------------- clxcorrelate_fft_vcf -- */
/* clxcorrelate_fft_vcf::make(fftSize, num_inputs, .., input_type=1) include/clenabled/
 *   clxcorrelate_fft_vcf.h:49; input_type 1 = spectra, 2 = time series (forward FFT first).
 * work (lib/clxcorrelate_fft_vcf_impl.cc:1057-1145): per vector and non-reference input k,
 *   out[k-1] = fftshift(|IFFT(ref * conj(in[k]))|), backward transform unscaled (:727).         */
#define CLB200_XCFFT_MAX_INPUTS 32
CLB200_API int clb200_xcorr_fft_create(int fft_size, int num_inputs, int input_type, int device,
                                       clb200_handle *out);
/* in: num_inputs pointers to nvec*fft_size c32; out: num_inputs-1 pointers to nvec*fft_size f32 */
CLB200_API int clb200_xcorr_fft_work(clb200_handle h, const void *const *in, void *const *out, long nvec);
CLB200_API int clb200_xcorr_fft_launch_device(clb200_handle h, const void *const *d_in, void *const *d_out,
                                              long nvec, void *stream);

/* -------------------------------------------------------- clComplexFilter -- */
/* clComplexFilter::make(.., decimation, const std::vector<gr_complex>& taps, ..)
 *   include/clenabled/clComplexFilter.h:706; kernel td_FIR_complex_complex
 *   (lib/clComplexFilter_impl.cc:805-829): out[g] = sum_i taps[K-1-i] * in[g+i] with complex taps.
 * Streaming semantics as clb200_filter_*: history and decimation phase live in the handle.     */
CLB200_API int clb200_cfilter_create(int device, int decimation, const float *taps_c32, int ntaps,
                                     clb200_handle *out);
CLB200_API int clb200_cfilter_set_taps(clb200_handle h, const float *taps_c32, int ntaps);
CLB200_API int clb200_cfilter_work(clb200_handle h, const void *in, long n_in, void *out, long *n_out);
CLB200_API int clb200_cfilter_launch_device(clb200_handle h, const void *d_in, long n_in, void *d_out,
                                            long *n_out, void *stream);

/* ------------------------------------------------------ clQuadratureDemod -- */
/* clQuadratureDemod::make(gain, ..) include/clenabled/clQuadratureDemod.h:49; kernel quadDemod
 *   (lib/clQuadratureDemod_impl.cc:118-143): out[i] = gain * atan2 of x[i+1] * conj(x[i]) in double;
 *   set_history(2) (:81): the previous call's last sample is kept in the handle (first call: 0).  */
CLB200_API int clb200_quaddemod_create(int device, float gain, clb200_handle *out);
CLB200_API int clb200_quaddemod_work(clb200_handle h, const void *in, void *out, long nitems);
CLB200_API int clb200_quaddemod_launch_device(clb200_handle h, const void *d_in, void *d_out, long nitems,
                                              void *stream);

/* --------------------------------------------------------- clSignalSource -- */
/* clSignalSource::make(idataType, .., samp_rate, waveform, freq, amplitude, ..)
 *   include/clenabled/clSignalSource.h:49-50; kernels sig_float / sig_complex, double branch
 *   (lib/clSignalSource_impl.cc:128-211): value(index) = f(phase + phase_inc*index) * ampl with
 *   phase_inc = 2*pi*freq/samp_rate; after each call the phase advances by phase_inc*n and whole
 *   turns are dropped (:386-398).  waveform: 1 cos, 2 sin (float output); complex = (cos, sin).  */
#define CLB200_SIG_COS 1   /* SIGSOURCE_COS, lib/clSignalSource_impl.h:27 */
#define CLB200_SIG_SIN 2
CLB200_API int clb200_sigsource_create(int device, int data_type, double samp_rate, int waveform,
                                       double freq, double amplitude, clb200_handle *out);
CLB200_API int clb200_sigsource_work(clb200_handle h, void *out, long nitems);
CLB200_API int clb200_sigsource_launch_device(clb200_handle h, void *d_out, long nitems, void *stream);
CLB200_API double clb200_sigsource_phase(clb200_handle h);

#ifdef __cplusplus
}
#endif
#endif /* CLENABLED_B200_H */
