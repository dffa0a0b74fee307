// Public block classes of the B200 build of gr-clenabled's hot path.
//
// Class names, namespace, base classes and make() signatures are the reference's
// (include/clenabled/clMathConst.h:51, clMathOp.h:42, clFFT.h:54-55 with the argument
// order of lib/clFFT_impl.cc:35-36, clFilter.h:52-53, clPolyphaseChannelizer.h:48-49,
// clXEngine.h:48-52, clLog.h / clSNR.h / clComplexToMag.h / clComplexToArg.h /
// clComplexToMagPhase.h / clMagPhaseToComplex.h), so a flowgraph or GRC-generated
// python file that instantiates them keeps working.  The four OpenCL selector ints
// are accepted unchanged and mapped to a CUDA ordinal (clb200_select_device).
// Constructor argument errors throw like the reference's constructors do; a runtime
// CUDA failure inside work() is logged and returns WORK_DONE instead of exit(0).
#ifndef INCLUDED_CLENABLED_B200_BLOCKS_H
#define INCLUDED_CLENABLED_B200_BLOCKS_H

#include <clenabled/api.h>
#include <gnuradio/block.h>
#include <gnuradio/sync_block.h>
#include <gnuradio/sync_decimator.h>

#include <memory>
#include <string>
#include <vector>

// include/clenabled/GRCLBase.h:57-70
#define DTYPE_COMPLEX 1
#define DTYPE_FLOAT 2
#define DTYPE_INT 3
#define DTYPE_SHORT 4
#define DTYPE_BYTE 5
#define DTYPE_PACKEDXY 6
#define OCLTYPE_GPU 1
#define OCLTYPE_ACCELERATOR 2
#define OCLTYPE_CPU 3
#define OCLTYPE_ANY 4
#define OCLDEVICESELECTOR_FIRST 1
#define OCLDEVICESELECTOR_SPECIFIC 2
// include/clenabled/clMathOpTypes.h:11-20
#define MATHOP_MULTIPLY 1
#define MATHOP_ADD 2
#define MATHOP_SUBTRACT 3
#define MATHOP_COMPLEX_CONJUGATE 4
#define MATHOP_MULTIPLY_CONJUGATE 5
#define MATHOP_EMPTY 255
#define MATHOP_EMPTY_W_COPY 254
// clFFT direction (grc/clenabled_clFFT.block.yml:37-41)
#define CLFFT_FORWARD_DIR (-1)
#define CLFFT_BACKWARD_DIR (1)

namespace gr {
namespace clenabled {

class CLENABLED_API clMathConst : virtual public gr::sync_block
{
public:
    typedef std::shared_ptr<clMathConst> sptr;
    static sptr make(int idataType, int openCLPlatformType, int devSelector, int platformId, int devId,
                     float fValue, int operatorType, int setDebug = 0);
    virtual float k() const = 0;
    virtual void set_k(float newValue) = 0;
};

class CLENABLED_API clMathOp : virtual public gr::sync_block
{
public:
    typedef std::shared_ptr<clMathOp> sptr;
    static sptr make(int idataType, int openCLPlatformType, int devSelector, int platformId, int devId,
                     int operatorType, int setDebug = 0);
};

#define CLB200_DECLARE_SIMPLE(NAME, ...)                                                          \
    class CLENABLED_API NAME : virtual public gr::sync_block                                      \
    {                                                                                             \
    public:                                                                                       \
        typedef std::shared_ptr<NAME> sptr;                                                       \
        static sptr make(__VA_ARGS__);                                                            \
    };
CLB200_DECLARE_SIMPLE(clLog, int openCLPlatformType, int devSelector, int platformId, int devId,
                      float nValue, float kValue, int setDebug = 0)
CLB200_DECLARE_SIMPLE(clSNR, int openCLPlatformType, int devSelector, int platformId, int devId,
                      float nValue, float kValue, int setDebug = 0)
CLB200_DECLARE_SIMPLE(clComplexToMag, int openCLPlatformType, int devSelector, int platformId, int devId,
                      int setDebug = 0)
CLB200_DECLARE_SIMPLE(clComplexToArg, int openCLPlatformType, int devSelector, int platformId, int devId,
                      int setDebug = 0)
CLB200_DECLARE_SIMPLE(clComplexToMagPhase, int openCLPlatformType, int devSelector, int platformId, int devId,
                      int setDebug = 0)
CLB200_DECLARE_SIMPLE(clMagPhaseToComplex, int openCLPlatformType, int devSelector, int platformId, int devId,
                      int setDebug = 0)
#undef CLB200_DECLARE_SIMPLE

class CLENABLED_API clFFT : virtual public gr::sync_block
{
public:
    typedef std::shared_ptr<clFFT> sptr;
    // item = one vector of fftSize elements; num_streams parallel ports
    static sptr make(int fftSize, int clFFTDir, const std::vector<float> &window, int idataType,
                     int openCLPlatformType, int devSelector, int platformId, int devId, int setDebug = 0,
                     int num_streams = 1, bool shift = false);
};

class CLENABLED_API clFilter : virtual public gr::sync_decimator
{
public:
    typedef std::shared_ptr<clFilter> sptr;
    static sptr make(int openclPlatform, int devSelector, int platformId, int devId, int decimation,
                     const std::vector<float> &taps, int nthreads = 1, int setDebug = 0, bool use_time = false);
    virtual void set_taps2(const std::vector<float> &taps) = 0;
    virtual std::vector<float> taps() const = 0;
    virtual void set_nthreads(int n) = 0;
    virtual int nthreads() const = 0;
};

class CLENABLED_API clPolyphaseChannelizer : virtual public gr::block
{
public:
    typedef std::shared_ptr<clPolyphaseChannelizer> sptr;
    static sptr make(int openCLPlatformType, int devSelector, int platformId, int devId,
                     const std::vector<float> &taps, int buf_items, int num_channels, int ninputs_per_iter,
                     const std::vector<int> &ch_map, int setDebug = 0);
};

class CLENABLED_API clXEngine : virtual public gr::block
{
public:
    typedef std::shared_ptr<clXEngine> sptr;
    static sptr make(int openCLPlatformType, int devSelector, int platformId, int devId, bool setDebug,
                     int data_type, int polarization, int num_inputs, int output_format, int first_channel,
                     int num_channels, int integration, std::vector<std::string> antenna_list,
                     bool output_file = false, std::string file_base = "", int rollover_size_mb = 0,
                     bool internal_synchronizer = false, long sync_timestamp = 0, std::string object_name = "",
                     double starting_chan_center_freq = 0.0, double channel_width = 0.0,
                     bool disable_output = false, int pipeline_integration = 0);
};

// ---- SURVEY 8(f) "next" rows ------------------------------------------------------------------
// include/clenabled/clXCorrelate.h:56-57; message port "corr" carries the (corrvect, corrective_lags) dict
class CLENABLED_API clXCorrelate : virtual public gr::sync_block
{
public:
    typedef std::shared_ptr<clXCorrelate> sptr;
    static sptr make(int openCLPlatformType, int devSelector, int platformId, int devId, bool setDebug,
                     int num_inputs, int signal_length, int data_type, int data_size, int max_search_index,
                     int decim_frames, bool async = false);
};

// include/clenabled/clxcorrelate_fft_vcf.h:49
class CLENABLED_API clxcorrelate_fft_vcf : virtual public gr::sync_block
{
public:
    typedef std::shared_ptr<clxcorrelate_fft_vcf> sptr;
    static sptr make(int fftSize, int num_inputs, int openCLPlatformType, int devSelector, int platformId,
                     int devId, int input_type = 1);
};

// include/clenabled/clComplexFilter.h:706
class CLENABLED_API clComplexFilter : virtual public gr::sync_decimator
{
public:
    typedef std::shared_ptr<clComplexFilter> sptr;
    static sptr make(int openclPlatform, int devSelector, int platformId, int devId, int decimation,
                     const std::vector<gr_complex> &taps, int nthreads = 1, int setDebug = 0);
    virtual void set_taps2(const std::vector<gr_complex> &taps) = 0;
    virtual std::vector<gr_complex> taps() const = 0;
};

// include/clenabled/clQuadratureDemod.h:49
class CLENABLED_API clQuadratureDemod : virtual public gr::sync_block
{
public:
    typedef std::shared_ptr<clQuadratureDemod> sptr;
    static sptr make(float gain, int openCLPlatformType, int devSelector, int platformId, int devId,
                     int setDebug = 0);
};

// include/clenabled/clSignalSource.h:49-50; waveform SIGSOURCE_COS 1 / SIGSOURCE_SIN 2
#define SIGSOURCE_COS 1
#define SIGSOURCE_SIN 2
class CLENABLED_API clSignalSource : virtual public gr::sync_block
{
public:
    typedef std::shared_ptr<clSignalSource> sptr;
    static sptr make(int idataType, int openCLPlatformType, int devSelector, int platformId, int devId,
                     double samp_rate, int waveform, double freq, float amplitude, int setDebug = 0);
};

} // namespace clenabled
} // namespace gr
#endif
