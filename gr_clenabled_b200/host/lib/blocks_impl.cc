// blocks_impl.cc -- the gr::clenabled block classes over the C ABI (clenabled_b200.h).
//
// Mirrors the reference's *_impl classes (lib/clMathConst_impl.cc, clMathOp_impl.cc,
// clFFT_impl.cc, clFilter_impl.cc, clPolyphaseChannelizer_impl.cc, clXEngine_impl.cc and
// the secondary element-wise blocks): same io signatures, scheduler hints (set_history,
// set_output_multiple), item semantics and error behaviour; all arithmetic happens in
// libclenabled_b200.so.  No OpenCL, no CPU path.
#include <clenabled/blocks.h>

#include <clenabled_b200.h>

#include <cstdio>
#include <cstring>
#include <stdexcept>

namespace gr {
namespace clenabled {

namespace {

int pick_device(int platformType, int devSelector, int platformId, int devId)
{
    int d = clb200_select_device(platformType, devSelector, platformId, devId);
    if (d < 0) throw std::runtime_error(std::string("clenabled_b200: ") + clb200_last_error());
    return d;
}

// ctor-time failures throw (reference ctors throw std::runtime_error / invalid_argument /
// out_of_range); EINVAL maps to invalid_argument
void must(int rc, bool out_of_range = false)
{
    if (rc == CLB200_OK) return;
    std::string msg = clb200_last_error();
    if (rc == CLB200_EINVAL) {
        if (out_of_range) throw std::out_of_range(msg);
        throw std::invalid_argument(msg);
    }
    throw std::runtime_error(msg);
}

// work-time failures: log and stop the block (reference: print + exit(0), GRCLBase.cpp:239-257)
int work_failed(const char *who)
{
    fprintf(stderr, "%s: %s\n", who, clb200_last_error());
    return gr::block::WORK_DONE;
}

struct Handle {
    clb200_handle h = nullptr;
    ~Handle() { clb200_destroy(h); }
};

size_t item_size(int idataType)
{
    return idataType == DTYPE_COMPLEX ? sizeof(gr_complex) : (idataType == DTYPE_FLOAT ? sizeof(float) : sizeof(int));
}

// ------------------------------------------------------------------ clMathConst --
class clMathConst_impl : public clMathConst
{
    Handle d;

public:
    clMathConst_impl(int idataType, int dev, float fValue, int operatorType)
        : gr::sync_block("clMathConst", gr::io_signature::make(1, 1, item_size(idataType)),
                         gr::io_signature::make(1, 1, item_size(idataType)))
    {
        must(clb200_mathconst_create(idataType, dev, fValue, operatorType, &d.h));
        set_output_multiple(256);       // reference: preferred work-group multiple (clMathConst_impl.cc:86-94)
    }
    float k() const override { return clb200_mathconst_k(d.h); }
    void set_k(float v) override { clb200_mathconst_set_k(d.h, v); }
    int work(int noutput_items, gr_vector_const_void_star &in, gr_vector_void_star &out) override
    {
        if (clb200_mathconst_work(d.h, in[0], out[0], noutput_items) != CLB200_OK)
            return work_failed("clMathConst");
        return noutput_items;
    }
};

// --------------------------------------------------------------------- clMathOp --
class clMathOp_impl : public clMathOp
{
    Handle d;

public:
    clMathOp_impl(int idataType, int dev, int operatorType)
        : gr::sync_block("clMathOp", gr::io_signature::make(2, 2, item_size(idataType)),
                         gr::io_signature::make(1, 1, item_size(idataType)))
    {
        must(clb200_mathop_create(idataType, dev, operatorType, &d.h));
        set_output_multiple(256);
    }
    int work(int noutput_items, gr_vector_const_void_star &in, gr_vector_void_star &out) override
    {
        if (noutput_items == 0) return 0;       // clMathOp_impl.cc:367-369
        if (clb200_mathop_work(d.h, in[0], in[1], out[0], noutput_items) != CLB200_OK)
            return work_failed("clMathOp");
        return noutput_items;
    }
};

// ------------------------------------------------------------ secondary blocks --
template <class Base>
class unary_impl : public Base
{
    Handle d;

public:
    unary_impl(const char *name, int kind, size_t in_size, int dev, float n, float k)
        : gr::sync_block(name, gr::io_signature::make(1, 1, in_size), gr::io_signature::make(1, 1, sizeof(float)))
    {
        must(clb200_unary_create(kind, dev, n, k, &d.h));
    }
    int work(int noutput_items, gr_vector_const_void_star &in, gr_vector_void_star &out) override
    {
        if (clb200_unary_work(d.h, in[0], out[0], noutput_items) != CLB200_OK) return work_failed("clenabled");
        return noutput_items;
    }
};

class clSNR_impl : public clSNR
{
    Handle d;

public:
    clSNR_impl(int dev, float n, float k)
        : gr::sync_block("clSNR", gr::io_signature::make(2, 2, sizeof(float)), gr::io_signature::make(1, 1, sizeof(float)))
    {
        must(clb200_snr_create(dev, n, k, &d.h));
    }
    int work(int n, gr_vector_const_void_star &in, gr_vector_void_star &out) override
    {
        if (clb200_snr_work(d.h, (const float *)in[0], (const float *)in[1], (float *)out[0], n) != CLB200_OK)
            return work_failed("clSNR");
        return n;
    }
};

class clComplexToMagPhase_impl : public clComplexToMagPhase
{
    Handle d;

public:
    explicit clComplexToMagPhase_impl(int dev)
        : gr::sync_block("clComplexToMagPhase", gr::io_signature::make(1, 1, sizeof(gr_complex)),
                         gr::io_signature::make(2, 2, sizeof(float)))
    {
        must(clb200_c2magphase_create(dev, &d.h));
    }
    int work(int n, gr_vector_const_void_star &in, gr_vector_void_star &out) override
    {
        if (clb200_c2magphase_work(d.h, in[0], (float *)out[0], (float *)out[1], n) != CLB200_OK)
            return work_failed("clComplexToMagPhase");
        return n;
    }
};

class clMagPhaseToComplex_impl : public clMagPhaseToComplex
{
    Handle d;

public:
    explicit clMagPhaseToComplex_impl(int dev)
        : gr::sync_block("clMagPhaseToComplex", gr::io_signature::make(2, 2, sizeof(float)),
                         gr::io_signature::make(1, 1, sizeof(gr_complex)))
    {
        must(clb200_magphase2c_create(dev, &d.h));
    }
    int work(int n, gr_vector_const_void_star &in, gr_vector_void_star &out) override
    {
        if (clb200_magphase2c_work(d.h, (const float *)in[0], (const float *)in[1], out[0], n) != CLB200_OK)
            return work_failed("clMagPhaseToComplex");
        return n;
    }
};

// ------------------------------------------------------------------------ clFFT --
class clFFT_impl : public clFFT
{
    Handle d;
    int d_fft_size, d_num_streams;

public:
    clFFT_impl(int fftSize, int dir, const std::vector<float> &window, int idataType, int dev, int num_streams,
               bool shift)
        : gr::sync_block("clFFT",
                         gr::io_signature::make(num_streams, num_streams, fftSize * item_size(idataType)),
                         gr::io_signature::make(num_streams, num_streams, fftSize * sizeof(gr_complex))),
          d_fft_size(fftSize), d_num_streams(num_streams)
    {
        // lib/clFFT_impl.cc:74-76
        if (!window.empty() && (int)window.size() != fftSize)
            throw std::runtime_error("fft_vcc: window not the same length as fft_size");
        if (num_streams < 1) throw std::invalid_argument("clFFT: num_streams must be >= 1");
        must(clb200_fft_create(fftSize, dir, window.empty() ? nullptr : window.data(), (int)window.size(),
                               idataType, dev, shift ? 1 : 0, &d.h));
    }
    // noutput_items counts vectors (lib/clFFT_impl.cc:637-654)
    int work(int noutput_items, gr_vector_const_void_star &in, gr_vector_void_star &out) override
    {
        if (clb200_fft_work_streams(d.h, in.data(), out.data(), d_num_streams, noutput_items) != CLB200_OK)
            return work_failed("clFFT");
        return noutput_items;
    }
};

// --------------------------------------------------------------------- clFilter --
class clFilter_impl : public clFilter
{
    Handle d;
    int d_nthreads;

public:
    clFilter_impl(int dev, int decimation, const std::vector<float> &taps, int nthreads, bool use_time)
        : gr::sync_decimator("clFilter", gr::io_signature::make(1, 1, sizeof(gr_complex)),
                             gr::io_signature::make(1, 1, sizeof(gr_complex)), decimation),
          d_nthreads(nthreads)
    {
        must(clb200_filter_create(dev, decimation, taps.data(), (int)taps.size(), use_time ? 1 : 0, &d.h));
        // The reference sets set_history(ntaps) (clFilter_impl.cc:78) because its kernels read
        // the overlap out of the scheduler's buffer.  Here the last ntaps-1 samples live in the
        // handle on the device, so the block consumes exactly noutput*decimation NEW samples
        // and in[0] is the first new one: history stays 1.
    }
    void set_taps2(const std::vector<float> &taps) override
    {
        if (clb200_filter_set_taps(d.h, taps.data(), (int)taps.size()) != CLB200_OK)
            throw std::invalid_argument(clb200_last_error());
    }
    std::vector<float> taps() const override
    {
        std::vector<float> t(clb200_filter_ntaps(d.h));
        clb200_filter_get_taps(d.h, t.data(), (int)t.size());
        return t;
    }
    void set_nthreads(int n) override { d_nthreads = n; }      // accepted, unused like the reference (:57)
    int nthreads() const override { return d_nthreads; }
    int work(int noutput_items, gr_vector_const_void_star &in, gr_vector_void_star &out) override
    {
        long n_out = 0;
        if (clb200_filter_work(d.h, in[0], (long)noutput_items * decimation(), out[0], &n_out) != CLB200_OK)
            return work_failed("clFilter");
        return (int)n_out;
    }
};

// ------------------------------------------------------- clPolyphaseChannelizer --
class clPolyphaseChannelizer_impl : public clPolyphaseChannelizer
{
    Handle d;
    int d_ntaps, d_buf_items, d_M, d_R, d_nmap;

public:
    clPolyphaseChannelizer_impl(int dev, const std::vector<float> &taps, int buf_items, int num_channels,
                                int ninputs_per_iter, const std::vector<int> &ch_map)
        : gr::block("clPolyphaseChannelizer", gr::io_signature::make(1, 1, sizeof(gr_complex)),
                    gr::io_signature::make(1, 1, sizeof(gr_complex))),
          d_ntaps((int)taps.size()), d_buf_items(buf_items), d_M(num_channels), d_R(ninputs_per_iter),
          d_nmap((int)ch_map.size())
    {
        must(clb200_pfb_create(dev, taps.data(), d_ntaps, buf_items, num_channels, ninputs_per_iter, ch_map.data(),
                               d_nmap, &d.h));
        set_history(d_ntaps);                                        // :63
        set_output_multiple(d_nmap * d_buf_items / d_R);             // :64
    }
    void forecast(int noutput_items, gr_vector_int &req) override
    {
        // one call = buf_items/R time steps, which read (steps-1)*R + ntaps samples; the
        // reference asks for R*nout/nmap + history - M (:77-81), which is short when R < M
        req[0] = d_R * noutput_items / d_nmap + (int)history() - d_R;
    }
    int general_work(int, gr_vector_int &, gr_vector_const_void_star &in, gr_vector_void_star &out) override
    {
        const long niter = d_buf_items / d_R;
        if (clb200_pfb_work(d.h, in[0], out[0], niter) != CLB200_OK) return work_failed("clPolyphaseChannelizer");
        consume_each(d_buf_items);                                   // :105
        return (int)(d_nmap * niter);                                // :108
    }
};

// -------------------------------------------------------------------- clXEngine --
class clXEngine_impl : public clXEngine
{
    Handle d;
    int d_data_type, d_npol, d_num_inputs, d_num_channels, d_integration, d_pipeline;
    bool d_disable_output;
    size_t d_sample_bytes;
    std::vector<char> d_buf;                  // one integration, [t][station][chan][pol]
    std::vector<gr_complex> d_matrix;
    int d_tracker = 0, d_pipeline_count = 0;
    // file sink (lib/clXEngine_impl.cc:393-465, :1259-1277): raw cf32_le frames + JSON sidecar
    bool d_output_file;
    std::string d_file_base, d_object_name, d_filename;
    std::vector<std::string> d_antennas;
    long d_rollover_bytes, d_bytes_written = 0, d_sync_timestamp, d_frames = 0;
    int d_rollover_index = 0, d_first_channel;
    double d_chan_freq, d_chan_width;
    FILE *d_fp = nullptr;
    bool d_use_sync, d_synchronized = false;      // ATA SNAP tag synchroniser (:1158-1226)
    uint64_t d_current_timestamp = 0;
    std::vector<uint64_t> d_tag_list;

    bool open_file()
    {
        if (d_fp) fclose(d_fp);
        d_filename = d_file_base;
        if (d_rollover_bytes > 0) {                         // "_NNN" suffix (:405-412)
            char suffix[16];
            snprintf(suffix, sizeof(suffix), "_%03d", d_rollover_index++);
            d_filename += suffix;
        }
        d_fp = fopen(d_filename.c_str(), "wb");
        d_bytes_written = 0;
        if (!d_fp) return false;
        // sidecar with the reference's keys (:438-465)
        FILE *js = fopen((d_filename + ".json").c_str(), "w");
        if (js) {
            std::string names = "[";
            for (size_t i = 0; i < d_antennas.size(); i++) names += (i ? ",\"" : "\"") + d_antennas[i] + "\"";
            names += "]";
            const long ntime = (long)d_integration * (d_pipeline > 1 ? d_pipeline : 1);
            fprintf(js,
                    "{\n\"sync_timestamp\":%ld,\n\"first_seq_num\":%ld,\n\"object_name\":\"%s\",\n"
                    "\"num_baselines\":%d,\n\"first_channel\":%d,\n\"first_channel_center_freq\":%f,\n"
                    "\"channels\":%d,\n\"channel_width\":%f,\n\"polarizations\":%d,\n\"antennas\":%d,\n"
                    "\"antenna_names\":%s,\n\"ntime\":%ld,\n\"samples_per_block\":%ld,\n"
                    "\"bytes_per_block\":%ld,\n\"data_type\":\"cf32_le\",\n\"data_format\": \"triangular order\"\n}\n",
                    d_sync_timestamp, d_frames * ntime, d_object_name.c_str(), d_num_inputs * (d_num_inputs + 1) / 2,
                    d_first_channel, d_chan_freq, d_num_channels, d_chan_width, d_npol, d_num_inputs, names.c_str(),
                    ntime, (long)d_matrix.size(), (long)(d_matrix.size() * sizeof(gr_complex)));
            fclose(js);
        }
        return true;
    }
    void write_frame()
    {
        const long bytes = (long)(d_matrix.size() * sizeof(gr_complex));
        if (!d_fp || (d_rollover_bytes > 0 && d_bytes_written + bytes > d_rollover_bytes))
            if (!open_file()) return;
        d_bytes_written += (long)fwrite(d_matrix.data(), 1, bytes, d_fp);
        d_frames++;
    }

public:
    clXEngine_impl(int dev, int data_type, int polarization, int num_inputs, int num_channels, int integration,
                   bool disable_output, int pipeline_integration, bool output_file, const std::string &file_base,
                   int rollover_size_mb, const std::vector<std::string> &antenna_list, long sync_timestamp,
                   const std::string &object_name, int first_channel, double chan_freq, double chan_width,
                   bool internal_synchronizer)
        : gr::block("clXEngine",
                    gr::io_signature::make(2, num_inputs * (data_type == DTYPE_PACKEDXY ? 1 : polarization),
                                           num_channels * (data_type == DTYPE_PACKEDXY ? 2
                                                           : data_type == DTYPE_BYTE   ? 2
                                                                                       : (int)sizeof(gr_complex))),
                    gr::io_signature::make(0, 0, 0)),
          d_data_type(data_type), d_npol(polarization), d_num_inputs(num_inputs), d_num_channels(num_channels),
          d_integration(integration), d_pipeline(pipeline_integration), d_disable_output(disable_output),
          d_output_file(output_file), d_file_base(file_base), d_object_name(object_name), d_antennas(antenna_list),
          d_rollover_bytes((long)rollover_size_mb * 1000000L), d_sync_timestamp(sync_timestamp),
          d_first_channel(first_channel), d_chan_freq(chan_freq), d_chan_width(chan_width),
          d_use_sync(internal_synchronizer)
    {
        if (internal_synchronizer && (integration % 16) > 0)          // :111-116
            throw std::out_of_range("ATA xengine: The number of integration frames should be a multiple of 16 to "
                                    "align with blocks coming from the SNAP.");
        must(clb200_xengine_create(dev, data_type, polarization, num_inputs, num_channels, integration, &d.h),
             num_inputs < 2);                   // std::out_of_range, clXEngine_impl.cc:106-109
        d_sample_bytes = data_type == DTYPE_COMPLEX ? sizeof(gr_complex) : (data_type == DTYPE_BYTE ? 2 : 1);
        d_buf.resize((size_t)clb200_xengine_input_bytes(d.h));
        d_matrix.resize((size_t)clb200_xengine_output_items(d.h));
        message_port_register_out(pmt::mp("xcorr"));                 // :294-295
        message_port_register_out(pmt::mp("sync"));
        if (d_use_sync) {                                            // :297-301
            set_tag_propagation_policy(TPP_DONT);
            set_output_multiple(16);
            d_tag_list.resize(num_inputs * (data_type == DTYPE_PACKEDXY ? 1 : polarization));
        }
    }
    ~clXEngine_impl() override
    {
        if (d_fp) fclose(d_fp);
    }
    bool stop() override
    {
        if (d_fp) {
            fclose(d_fp);
            d_fp = nullptr;
        }
        return true;
    }
    void forecast(int noutput_items, gr_vector_int &req) override
    {
        for (auto &r : req) r = noutput_items;                       // :385-390
    }
    // work_processor (:918-1142): marshal the port vectors of each time step into the
    // integration buffer; a full buffer is correlated and published as ("triang_matrix" . c32vector)
    int general_work(int noutput_items, gr_vector_int &, gr_vector_const_void_star &in, gr_vector_void_star &) override
    {
        gr::thread::scoped_lock guard(d_setlock);
        if (d_use_sync && !d_synchronized) {
            // SNAP packets carry a sequence tag on every 16-step block (t[n+1] = t[n] + 16).  Until the first
            // tag of every input is the same, drop (highest - own) items from each input and produce nothing.
            uint64_t highest = 0, first = 0;
            bool in_sync = true;
            for (size_t p = 0; p < d_tag_list.size(); p++) {
                std::vector<gr::tag_t> tags;
                get_tags_in_window(tags, (unsigned)p, 0, 1);
                if (tags.empty()) return 0;                    // (the reference would dereference tags[0])
                const uint64_t t0 = pmt::to_uint64(tags[0].value);
                if (p == 0) first = t0;
                else if (t0 != first) in_sync = false;
                d_tag_list[p] = t0;
                highest = std::max(highest, t0);
            }
            if (!in_sync) {
                for (size_t p = 0; p < d_tag_list.size(); p++)
                    consume((int)p, (int)std::min<uint64_t>(highest - d_tag_list[p], (uint64_t)noutput_items));
                return 0;
            }
            d_synchronized = true;
            d_current_timestamp = highest;
            d_sync_timestamp = (long)highest;                   // goes into the JSON sidecar (write_json(highest_tag))
            message_port_pub(pmt::mp("sync"), pmt::cons(pmt::intern("synctimestamp"), pmt::from_uint64(highest)));
        }
        int n = std::min(noutput_items, d_integration - d_tracker);
        const size_t vec = (size_t)d_num_channels * d_sample_bytes;          // bytes of one port item
        const size_t row = vec * d_npol;                                      // one station, one time step
        const size_t frame = row * d_num_inputs;
        for (int t = 0; t < n; t++) {
            char *dst = d_buf.data() + (size_t)(d_tracker + t) * frame;
            for (int s = 0; s < d_num_inputs; s++) {
                if (d_npol == 1 || d_data_type == DTYPE_PACKEDXY) {
                    memcpy(dst + s * row, (const char *)in[s] + (size_t)t * row, row);
                } else {                                                      // interleave X,Y per channel (:1010-1057)
                    const char *x = (const char *)in[s] + (size_t)t * vec;
                    const char *y = (const char *)in[s + d_num_inputs] + (size_t)t * vec;
                    char *o = dst + s * row;
                    for (int c = 0; c < d_num_channels; c++) {
                        memcpy(o + (2 * c) * d_sample_bytes, x + c * d_sample_bytes, d_sample_bytes);
                        memcpy(o + (2 * c + 1) * d_sample_bytes, y + c * d_sample_bytes, d_sample_bytes);
                    }
                }
            }
        }
        d_tracker += n;
        if (d_tracker == d_integration) {
            const bool accumulate = d_pipeline > 1 && d_pipeline_count > 0;   // :785-808
            if (clb200_xengine_work(d.h, d_buf.data(), d_matrix.data(), accumulate ? 1 : 0) != CLB200_OK)
                return work_failed("clXEngine");
            d_tracker = 0;
            d_pipeline_count++;
            if (d_pipeline < 2 || d_pipeline_count >= d_pipeline) {
                d_pipeline_count = 0;
                if (d_output_file) write_frame();                    // :1259-1277
                else if (!d_disable_output)
                    message_port_pub(pmt::mp("xcorr"),
                                     pmt::cons(pmt::string_to_symbol("triang_matrix"),
                                               pmt::init_c32vector(d_matrix.size(), d_matrix.data())));   // :1076-1080
            }
        }
        for (size_t p = 0; p < in.size(); p++) consume((int)p, n);            // :1228-1231
        return n;
    }
};

// ------------------------------------------------------------------ clXCorrelate --
// lib/clXCorrelate_impl.cc:701-842 (ctor), :1528-1594 (synchronous work path)
class clXCorrelate_impl : public clXCorrelate
{
    Handle d;
    int d_num_inputs, d_signal_length, d_decim_frames, d_frame = 1;
    std::vector<float> d_corr;
    std::vector<int32_t> d_lag;

public:
    clXCorrelate_impl(int dev, int num_inputs, int signal_length, int data_type, int data_size, int max_search_index,
                      int decim_frames)
        : gr::sync_block("clXCorrelate", gr::io_signature::make(2, num_inputs, data_size),
                         gr::io_signature::make(0, 0, 0)),
          d_num_inputs(num_inputs), d_signal_length(signal_length), d_decim_frames(decim_frames)
    {
        if (data_size == 0) throw std::invalid_argument("clXCorrelate: Unknown data type.");     // :710-714
        must(clb200_xcorrelate_create(dev, num_inputs, signal_length, data_type, max_search_index, &d.h));
        d_corr.resize(num_inputs - 1);
        d_lag.resize(num_inputs - 1);
        set_output_multiple(signal_length);                                                       // :839
        message_port_register_out(pmt::mp("corr"));                                               // :840
    }
    int work(int noutput_items, gr_vector_const_void_star &in, gr_vector_void_star &) override
    {
        if (noutput_items < d_signal_length) return 0;                                            // :1532-1534
        if (d_decim_frames > 1) {                                                                 // :1539-1547
            if ((d_frame++ % d_decim_frames) == 0) d_frame = 1;
            else return d_signal_length;
        }
        if (clb200_xcorrelate_work(d.h, in.data(), d_corr.data(), d_lag.data()) != CLB200_OK)
            return work_failed("clXCorrelate");
        pmt::pmt_t meta = pmt::make_dict();                                                       // :1585-1593
        meta = pmt::dict_add(meta, pmt::mp("corrvect"), pmt::init_f32vector(d_corr.size(), d_corr.data()));
        meta = pmt::dict_add(meta, pmt::mp("corrective_lags"), pmt::init_s32vector(d_lag.size(), d_lag.data()));
        message_port_pub(pmt::mp("corr"), pmt::cons(meta, pmt::PMT_NIL));
        return d_signal_length;
    }
};

// ----------------------------------------------------------- clxcorrelate_fft_vcf --
// lib/clxcorrelate_fft_vcf_impl.cc:699-750, :1057-1145: items are whole vectors of fftSize
class clxcorrelate_fft_vcf_impl : public clxcorrelate_fft_vcf
{
    Handle d;

public:
    clxcorrelate_fft_vcf_impl(int dev, int fftSize, int num_inputs, int input_type)
        : gr::sync_block("clxcorrelate_fft_vcf", gr::io_signature::make(2, num_inputs, sizeof(gr_complex) * fftSize),
                         gr::io_signature::make(1, num_inputs - 1, sizeof(float) * fftSize))
    {
        must(clb200_xcorr_fft_create(fftSize, num_inputs, input_type, dev, &d.h));
    }
    int work(int noutput_items, gr_vector_const_void_star &in, gr_vector_void_star &out) override
    {
        if (clb200_xcorr_fft_work(d.h, in.data(), out.data(), noutput_items) != CLB200_OK)
            return work_failed("clxcorrelate_fft_vcf");
        return noutput_items;
    }
};

// --------------------------------------------------------------- clComplexFilter --
class clComplexFilter_impl : public clComplexFilter
{
    Handle d;

public:
    clComplexFilter_impl(int dev, int decimation, const std::vector<gr_complex> &taps)
        : gr::sync_decimator("clComplexFilter", gr::io_signature::make(1, 1, sizeof(gr_complex)),
                             gr::io_signature::make(1, 1, sizeof(gr_complex)), decimation)
    {
        must(clb200_cfilter_create(dev, decimation, reinterpret_cast<const float *>(taps.data()), (int)taps.size(), &d.h));
        // history lives in the handle (see clFilter_impl above): in[0] is the first NEW sample
    }
    void set_taps2(const std::vector<gr_complex> &taps) override
    {
        if (clb200_cfilter_set_taps(d.h, reinterpret_cast<const float *>(taps.data()), (int)taps.size()) != CLB200_OK)
            throw std::invalid_argument(clb200_last_error());
    }
    int work(int noutput_items, gr_vector_const_void_star &in, gr_vector_void_star &out) override
    {
        long n_out = 0;
        if (clb200_cfilter_work(d.h, in[0], (long)noutput_items * decimation(), out[0], &n_out) != CLB200_OK)
            return work_failed("clComplexFilter");
        return (int)n_out;
    }
};

// ------------------------------------------------------------- clQuadratureDemod --
class clQuadratureDemod_impl : public clQuadratureDemod
{
    Handle d;

public:
    clQuadratureDemod_impl(int dev, float gain)
        : gr::sync_block("clQuadratureDemod", gr::io_signature::make(1, 1, sizeof(gr_complex)),
                         gr::io_signature::make(1, 1, sizeof(float)))
    {
        must(clb200_quaddemod_create(dev, gain, &d.h));
        set_output_multiple(32);        // lib/clQuadratureDemod_impl.cc:71; the previous sample lives in the handle
    }
    int work(int noutput_items, gr_vector_const_void_star &in, gr_vector_void_star &out) override
    {
        if (clb200_quaddemod_work(d.h, in[0], out[0], noutput_items) != CLB200_OK)
            return work_failed("clQuadratureDemod");
        return noutput_items;
    }
};

// ---------------------------------------------------------------- clSignalSource --
class clSignalSource_impl : public clSignalSource
{
    Handle d;

public:
    clSignalSource_impl(int dev, int idataType, double samp_rate, int waveform, double freq, float amplitude)
        : gr::sync_block("clSignalSource", gr::io_signature::make(0, 0, 0),
                         gr::io_signature::make(1, 1, item_size(idataType)))
    {
        must(clb200_sigsource_create(dev, idataType, samp_rate, waveform, freq, amplitude, &d.h));
    }
    int work(int noutput_items, gr_vector_const_void_star &, gr_vector_void_star &out) override
    {
        if (clb200_sigsource_work(d.h, out[0], noutput_items) != CLB200_OK) return work_failed("clSignalSource");
        return noutput_items;
    }
};

} // namespace

// ---------------------------------------------------------------- factories --
clMathConst::sptr clMathConst::make(int idataType, int plat, int sel, int pid, int did, float fValue,
                                    int operatorType, int)
{
    return gnuradio::get_initial_sptr(new clMathConst_impl(idataType, pick_device(plat, sel, pid, did), fValue, operatorType));
}
clMathOp::sptr clMathOp::make(int idataType, int plat, int sel, int pid, int did, int operatorType, int)
{
    return gnuradio::get_initial_sptr(new clMathOp_impl(idataType, pick_device(plat, sel, pid, did), operatorType));
}
clLog::sptr clLog::make(int plat, int sel, int pid, int did, float n, float k, int)
{
    return gnuradio::get_initial_sptr(new unary_impl<clLog>("clLog", CLB200_UNARY_LOG10, sizeof(float),
                                                            pick_device(plat, sel, pid, did), n, k));
}
clSNR::sptr clSNR::make(int plat, int sel, int pid, int did, float n, float k, int)
{
    return gnuradio::get_initial_sptr(new clSNR_impl(pick_device(plat, sel, pid, did), n, k));
}
clComplexToMag::sptr clComplexToMag::make(int plat, int sel, int pid, int did, int)
{
    return gnuradio::get_initial_sptr(new unary_impl<clComplexToMag>("clComplexToMag", CLB200_UNARY_COMPLEX_TO_MAG,
                                                                     sizeof(gr_complex), pick_device(plat, sel, pid, did), 0, 0));
}
clComplexToArg::sptr clComplexToArg::make(int plat, int sel, int pid, int did, int)
{
    return gnuradio::get_initial_sptr(new unary_impl<clComplexToArg>("clComplexToArg", CLB200_UNARY_COMPLEX_TO_ARG,
                                                                     sizeof(gr_complex), pick_device(plat, sel, pid, did), 0, 0));
}
clComplexToMagPhase::sptr clComplexToMagPhase::make(int plat, int sel, int pid, int did, int)
{
    return gnuradio::get_initial_sptr(new clComplexToMagPhase_impl(pick_device(plat, sel, pid, did)));
}
clMagPhaseToComplex::sptr clMagPhaseToComplex::make(int plat, int sel, int pid, int did, int)
{
    return gnuradio::get_initial_sptr(new clMagPhaseToComplex_impl(pick_device(plat, sel, pid, did)));
}
clFFT::sptr clFFT::make(int fftSize, int dir, const std::vector<float> &window, int idataType, int plat, int sel,
                        int pid, int did, int, int num_streams, bool shift)
{
    return gnuradio::get_initial_sptr(
        new clFFT_impl(fftSize, dir, window, idataType, pick_device(plat, sel, pid, did), num_streams, shift));
}
clFilter::sptr clFilter::make(int plat, int sel, int pid, int did, int decimation, const std::vector<float> &taps,
                              int nthreads, int, bool use_time)
{
    return gnuradio::get_initial_sptr(new clFilter_impl(pick_device(plat, sel, pid, did), decimation, taps, nthreads, use_time));
}
clPolyphaseChannelizer::sptr clPolyphaseChannelizer::make(int plat, int sel, int pid, int did,
                                                          const std::vector<float> &taps, int buf_items,
                                                          int num_channels, int ninputs_per_iter,
                                                          const std::vector<int> &ch_map, int)
{
    return gnuradio::get_initial_sptr(new clPolyphaseChannelizer_impl(pick_device(plat, sel, pid, did), taps, buf_items,
                                                                      num_channels, ninputs_per_iter, ch_map));
}
clXEngine::sptr clXEngine::make(int plat, int sel, int pid, int did, bool, int data_type, int polarization,
                                int num_inputs, int, int first_channel, int num_channels, int integration,
                                std::vector<std::string> antenna_list, bool output_file, std::string file_base,
                                int rollover_size_mb, bool internal_synchronizer, long sync_timestamp, std::string object_name,
                                double starting_chan_center_freq, double channel_width, bool disable_output,
                                int pipeline_integration)
{
    return gnuradio::get_initial_sptr(new clXEngine_impl(
        pick_device(plat, sel, pid, did), data_type, polarization, num_inputs, num_channels, integration,
        disable_output, pipeline_integration, output_file, file_base, rollover_size_mb, antenna_list, sync_timestamp,
        object_name, first_channel, starting_chan_center_freq, channel_width, internal_synchronizer));
}

clXCorrelate::sptr clXCorrelate::make(int plat, int sel, int pid, int did, bool, int num_inputs, int signal_length,
                                      int data_type, int data_size, int max_search_index, int decim_frames, bool)
{
    return gnuradio::get_initial_sptr(new clXCorrelate_impl(pick_device(plat, sel, pid, did), num_inputs, signal_length,
                                                            data_type, data_size, max_search_index, decim_frames));
}
clxcorrelate_fft_vcf::sptr clxcorrelate_fft_vcf::make(int fftSize, int num_inputs, int plat, int sel, int pid, int did,
                                                      int input_type)
{
    return gnuradio::get_initial_sptr(
        new clxcorrelate_fft_vcf_impl(pick_device(plat, sel, pid, did), fftSize, num_inputs, input_type));
}
clComplexFilter::sptr clComplexFilter::make(int plat, int sel, int pid, int did, int decimation,
                                            const std::vector<gr_complex> &taps, int, int)
{
    return gnuradio::get_initial_sptr(new clComplexFilter_impl(pick_device(plat, sel, pid, did), decimation, taps));
}
clQuadratureDemod::sptr clQuadratureDemod::make(float gain, int plat, int sel, int pid, int did, int)
{
    return gnuradio::get_initial_sptr(new clQuadratureDemod_impl(pick_device(plat, sel, pid, did), gain));
}
clSignalSource::sptr clSignalSource::make(int idataType, int plat, int sel, int pid, int did, double samp_rate,
                                          int waveform, double freq, float amplitude, int)
{
    return gnuradio::get_initial_sptr(
        new clSignalSource_impl(pick_device(plat, sel, pid, did), idataType, samp_rate, waveform, freq, amplitude));
}

} // namespace clenabled
} // namespace gr
