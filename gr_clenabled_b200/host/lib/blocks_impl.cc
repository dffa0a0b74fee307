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

// work-time failures: log through the block's logger and stop the block (reference: print + exit(0),
// GRCLBase.cpp:239-257)
int work_failed(gr::logger_ptr log, const char *who)
{
    GR_LOG_ERROR(log, std::string(who) + ": " + clb200_last_error());
    return gr::block::WORK_DONE;
}

// setDebug of the reference factories: log the device and kernel the block was configured with (the
// reference prints its OpenCL device and kernel source, e.g. clXEngine_impl.cc:700-703) and have the
// library report every work() call
void apply_debug(gr::logger_ptr log, clb200_handle h, int setDebug)
{
    if (!setDebug) return;
    char buf[768];
    if (clb200_describe(h, buf, (int)sizeof(buf)) == CLB200_OK) GR_LOG_INFO(log, std::string(buf));
    clb200_set_debug(h, 1);
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
    void enable_debug(int v) { apply_debug(this->d_logger, d.h, v); }

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
            return work_failed(this->d_logger, "clMathConst");
        return noutput_items;
    }
};

// --------------------------------------------------------------------- clMathOp --
class clMathOp_impl : public clMathOp
{
    Handle d;

public:
    void enable_debug(int v) { apply_debug(this->d_logger, d.h, v); }

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
            return work_failed(this->d_logger, "clMathOp");
        return noutput_items;
    }
};

// ------------------------------------------------------------ secondary blocks --
template <class Base>
class unary_impl : public Base
{
    Handle d;

public:
    void enable_debug(int v) { apply_debug(this->d_logger, d.h, v); }

    unary_impl(const char *name, int kind, size_t in_size, int dev, float n, float k)
        : gr::sync_block(name, gr::io_signature::make(1, 1, in_size), gr::io_signature::make(1, 1, sizeof(float)))
    {
        must(clb200_unary_create(kind, dev, n, k, &d.h));
    }
    int work(int noutput_items, gr_vector_const_void_star &in, gr_vector_void_star &out) override
    {
        if (clb200_unary_work(d.h, in[0], out[0], noutput_items) != CLB200_OK) return work_failed(this->d_logger, "clenabled");
        return noutput_items;
    }
};

class clSNR_impl : public clSNR
{
    Handle d;

public:
    void enable_debug(int v) { apply_debug(this->d_logger, d.h, v); }

    clSNR_impl(int dev, float n, float k)
        : gr::sync_block("clSNR", gr::io_signature::make(2, 2, sizeof(float)), gr::io_signature::make(1, 1, sizeof(float)))
    {
        must(clb200_snr_create(dev, n, k, &d.h));
    }
    int work(int n, gr_vector_const_void_star &in, gr_vector_void_star &out) override
    {
        if (clb200_snr_work(d.h, (const float *)in[0], (const float *)in[1], (float *)out[0], n) != CLB200_OK)
            return work_failed(this->d_logger, "clSNR");
        return n;
    }
};

class clComplexToMagPhase_impl : public clComplexToMagPhase
{
    Handle d;

public:
    void enable_debug(int v) { apply_debug(this->d_logger, d.h, v); }

    explicit clComplexToMagPhase_impl(int dev)
        : gr::sync_block("clComplexToMagPhase", gr::io_signature::make(1, 1, sizeof(gr_complex)),
                         gr::io_signature::make(2, 2, sizeof(float)))
    {
        must(clb200_c2magphase_create(dev, &d.h));
    }
    int work(int n, gr_vector_const_void_star &in, gr_vector_void_star &out) override
    {
        if (clb200_c2magphase_work(d.h, in[0], (float *)out[0], (float *)out[1], n) != CLB200_OK)
            return work_failed(this->d_logger, "clComplexToMagPhase");
        return n;
    }
};

class clMagPhaseToComplex_impl : public clMagPhaseToComplex
{
    Handle d;

public:
    void enable_debug(int v) { apply_debug(this->d_logger, d.h, v); }

    explicit clMagPhaseToComplex_impl(int dev)
        : gr::sync_block("clMagPhaseToComplex", gr::io_signature::make(2, 2, sizeof(float)),
                         gr::io_signature::make(1, 1, sizeof(gr_complex)))
    {
        must(clb200_magphase2c_create(dev, &d.h));
    }
    int work(int n, gr_vector_const_void_star &in, gr_vector_void_star &out) override
    {
        if (clb200_magphase2c_work(d.h, (const float *)in[0], (const float *)in[1], out[0], n) != CLB200_OK)
            return work_failed(this->d_logger, "clMagPhaseToComplex");
        return n;
    }
};

// ------------------------------------------------------------------------ clFFT --
class clFFT_impl : public clFFT
{
    Handle d;

public:
    void enable_debug(int v) { apply_debug(this->d_logger, d.h, v); }

private:
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
            return work_failed(this->d_logger, "clFFT");
        return noutput_items;
    }
};

// --------------------------------------------------------------------- clFilter --
class clFilter_impl : public clFilter
{
    Handle d;

public:
    void enable_debug(int v) { apply_debug(this->d_logger, d.h, v); }

private:
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
            return work_failed(this->d_logger, "clFilter");
        return (int)n_out;
    }
};

// ------------------------------------------------------- clPolyphaseChannelizer --
class clPolyphaseChannelizer_impl : public clPolyphaseChannelizer
{
    Handle d;

public:
    void enable_debug(int v) { apply_debug(this->d_logger, d.h, v); }

private:
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
        if (clb200_pfb_work(d.h, in[0], out[0], niter) != CLB200_OK) return work_failed(this->d_logger, "clPolyphaseChannelizer");
        consume_each(d_buf_items);                                   // :105
        return (int)(d_nmap * niter);                                // :108
    }
};

// -------------------------------------------------------------------- clXEngine --
class clXEngine_impl : public clXEngine
{
    Handle d;

public:
    void enable_debug(int v) { apply_debug(this->d_logger, d.h, v); }

private:
    int d_data_type, d_npol, d_num_inputs, d_num_channels, d_integration, d_pipeline;
    bool d_disable_output;
    static constexpr int RESULT_SLOTS = 4;
    size_t d_nports;
    std::vector<gr_complex> d_matrix;         // the matrix being delivered (file frame / PDU payload)
    int d_tracker = 0;                        // time steps of the current integration consumed so far
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
        d_nports = (size_t)num_inputs * (data_type == DTYPE_PACKEDXY ? 1 : polarization);
        // pinned double buffers, copy / compute / read-back streams, result ring (the reference's two pinned host
        // buffers + worker thread, :304-382, :1234-1299); pipeline_integration sums on the device (:785-808)
        must(clb200_xengine_stream_begin(d.h, pipeline_integration, RESULT_SLOTS));
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
        {   // everything handed to the GPU is delivered before the block stops (the reference joins its worker, :498-520)
            gr::thread::scoped_lock guard(d_setlock);
            if (d.h && !pickup(1)) GR_LOG_ERROR(this->d_logger, std::string("clXEngine: ") + clb200_last_error());
        }
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
    // a finished visibility matrix goes to the file sink or out as ("triang_matrix" . c32vector) (:1076-1080, :1259-1277)
    void deliver()
    {
        if (d_output_file) write_frame();
        else if (!d_disable_output)
            message_port_pub(pmt::mp("xcorr"), pmt::cons(pmt::string_to_symbol("triang_matrix"),
                                                         pmt::init_c32vector(d_matrix.size(), d_matrix.data())));
    }
    // pick up whatever the GPU has finished; wait != 0 drains everything in flight
    bool pickup(int wait)
    {
        for (;;) {
            int ready = 0;
            if (clb200_xengine_poll_result(d.h, d_matrix.data(), wait, &ready) != CLB200_OK) return false;
            if (!ready) return true;
            deliver();
        }
    }

public:
    // work_processor (:918-1142) + runThread (:1234-1299).  The reference marshals each call into one of two pinned
    // host buffers and a worker thread uploads + correlates a full buffer; the finished matrix is published when the
    // NEXT integration completes.  Here each call's time steps go to the device at once (clb200_xengine_push_timesteps:
    // pinned staging + copy stream), the correlation and the read-back of a full integration are enqueued without
    // waiting, and finished matrices are picked up on the following calls (or in stop()): the scheduler thread never
    // waits for a kernel.
    int general_work(int noutput_items, gr_vector_int &, gr_vector_const_void_star &in, gr_vector_void_star &) override
    {
        gr::thread::scoped_lock guard(d_setlock);
        if (!pickup(0)) return work_failed(this->d_logger, "clXEngine");
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
        const int n = std::min(noutput_items, d_integration - d_tracker);      // :924-933: never past the integration
        if (d_tracker + n == d_integration) {
            // this call completes an integration: keep the result ring from filling up (the reference holds the
            // scheduler on d_thread_active_lock here when its worker is still busy, :935-947)
            long pending = 0;
            clb200_xengine_stream_state(d.h, nullptr, nullptr, &pending, nullptr, nullptr);
            if (pending >= RESULT_SLOTS - 1 && !pickup(1)) return work_failed(this->d_logger, "clXEngine");
        }
        if (clb200_xengine_push_timesteps(d.h, in.data(), (int)d_nports, n) != CLB200_OK)
            return work_failed(this->d_logger, "clXEngine");
        d_tracker = (d_tracker + n) % d_integration;
        for (size_t p = 0; p < in.size(); p++) consume((int)p, n);            // :1228-1231
        return n;
    }
};

// ------------------------------------------------------------------ clXCorrelate --
// lib/clXCorrelate_impl.cc:701-842 (ctor), :1528-1594 (synchronous work path)
class clXCorrelate_impl : public clXCorrelate
{
    Handle d;

public:
    void enable_debug(int v) { apply_debug(this->d_logger, d.h, v); }

private:
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
            return work_failed(this->d_logger, "clXCorrelate");
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
    void enable_debug(int v) { apply_debug(this->d_logger, d.h, v); }

    clxcorrelate_fft_vcf_impl(int dev, int fftSize, int num_inputs, int input_type)
        : gr::sync_block("clxcorrelate_fft_vcf", gr::io_signature::make(2, num_inputs, sizeof(gr_complex) * fftSize),
                         gr::io_signature::make(1, num_inputs - 1, sizeof(float) * fftSize))
    {
        must(clb200_xcorr_fft_create(fftSize, num_inputs, input_type, dev, &d.h));
    }
    int work(int noutput_items, gr_vector_const_void_star &in, gr_vector_void_star &out) override
    {
        if (clb200_xcorr_fft_work(d.h, in.data(), out.data(), noutput_items) != CLB200_OK)
            return work_failed(this->d_logger, "clxcorrelate_fft_vcf");
        return noutput_items;
    }
};

// --------------------------------------------------------------- clComplexFilter --
class clComplexFilter_impl : public clComplexFilter
{
    std::vector<gr_complex> d_taps;
    Handle d;

public:
    void enable_debug(int v) { apply_debug(this->d_logger, d.h, v); }

    clComplexFilter_impl(int dev, int decimation, const std::vector<gr_complex> &taps)
        : gr::sync_decimator("clComplexFilter", gr::io_signature::make(1, 1, sizeof(gr_complex)),
                             gr::io_signature::make(1, 1, sizeof(gr_complex)), decimation),
          d_taps(taps)
    {
        must(clb200_cfilter_create(dev, decimation, reinterpret_cast<const float *>(taps.data()), (int)taps.size(), &d.h));
        // history lives in the handle (see clFilter_impl above): in[0] is the first NEW sample
    }
    void set_taps2(const std::vector<gr_complex> &taps) override
    {
        if (clb200_cfilter_set_taps(d.h, reinterpret_cast<const float *>(taps.data()), (int)taps.size()) != CLB200_OK)
            throw std::invalid_argument(clb200_last_error());
        d_taps = taps;
    }
    std::vector<gr_complex> taps() const override { return d_taps; }
    int work(int noutput_items, gr_vector_const_void_star &in, gr_vector_void_star &out) override
    {
        long n_out = 0;
        if (clb200_cfilter_work(d.h, in[0], (long)noutput_items * decimation(), out[0], &n_out) != CLB200_OK)
            return work_failed(this->d_logger, "clComplexFilter");
        return (int)n_out;
    }
};

// ------------------------------------------------------------- clQuadratureDemod --
class clQuadratureDemod_impl : public clQuadratureDemod
{
    Handle d;

public:
    void enable_debug(int v) { apply_debug(this->d_logger, d.h, v); }

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
            return work_failed(this->d_logger, "clQuadratureDemod");
        return noutput_items;
    }
};

// ---------------------------------------------------------------- clSignalSource --
class clSignalSource_impl : public clSignalSource
{
    Handle d;

public:
    void enable_debug(int v) { apply_debug(this->d_logger, d.h, v); }

    clSignalSource_impl(int dev, int idataType, double samp_rate, int waveform, double freq, float amplitude)
        : gr::sync_block("clSignalSource", gr::io_signature::make(0, 0, 0),
                         gr::io_signature::make(1, 1, item_size(idataType)))
    {
        must(clb200_sigsource_create(dev, idataType, samp_rate, waveform, freq, amplitude, &d.h));
    }
    int work(int noutput_items, gr_vector_const_void_star &, gr_vector_void_star &out) override
    {
        if (clb200_sigsource_work(d.h, out[0], noutput_items) != CLB200_OK) return work_failed(this->d_logger, "clSignalSource");
        return noutput_items;
    }
};

} // namespace

// ---------------------------------------------------------------- factories --
template <class Impl>
std::shared_ptr<Impl> finish(Impl *p, int setDebug)
{
    p->enable_debug(setDebug);
    return gnuradio::get_initial_sptr(p);
}

clMathConst::sptr clMathConst::make(int idataType, int plat, int sel, int pid, int did, float fValue,
                                    int operatorType, int setDebug)
{
    return finish(new clMathConst_impl(idataType, pick_device(plat, sel, pid, did), fValue, operatorType), setDebug);
}
clMathOp::sptr clMathOp::make(int idataType, int plat, int sel, int pid, int did, int operatorType, int setDebug)
{
    return finish(new clMathOp_impl(idataType, pick_device(plat, sel, pid, did), operatorType), setDebug);
}
clLog::sptr clLog::make(int plat, int sel, int pid, int did, float n, float k, int setDebug)
{
    return finish(new unary_impl<clLog>("clLog", CLB200_UNARY_LOG10, sizeof(float), pick_device(plat, sel, pid, did), n, k),
                  setDebug);
}
clSNR::sptr clSNR::make(int plat, int sel, int pid, int did, float n, float k, int setDebug)
{
    return finish(new clSNR_impl(pick_device(plat, sel, pid, did), n, k), setDebug);
}
clComplexToMag::sptr clComplexToMag::make(int plat, int sel, int pid, int did, int setDebug)
{
    return finish(new unary_impl<clComplexToMag>("clComplexToMag", CLB200_UNARY_COMPLEX_TO_MAG, sizeof(gr_complex),
                                                 pick_device(plat, sel, pid, did), 0, 0),
                  setDebug);
}
clComplexToArg::sptr clComplexToArg::make(int plat, int sel, int pid, int did, int setDebug)
{
    return finish(new unary_impl<clComplexToArg>("clComplexToArg", CLB200_UNARY_COMPLEX_TO_ARG, sizeof(gr_complex),
                                                 pick_device(plat, sel, pid, did), 0, 0),
                  setDebug);
}
clComplexToMagPhase::sptr clComplexToMagPhase::make(int plat, int sel, int pid, int did, int setDebug)
{
    return finish(new clComplexToMagPhase_impl(pick_device(plat, sel, pid, did)), setDebug);
}
clMagPhaseToComplex::sptr clMagPhaseToComplex::make(int plat, int sel, int pid, int did, int setDebug)
{
    return finish(new clMagPhaseToComplex_impl(pick_device(plat, sel, pid, did)), setDebug);
}
clFFT::sptr clFFT::make(int fftSize, int dir, const std::vector<float> &window, int idataType, int plat, int sel,
                        int pid, int did, int setDebug, int num_streams, bool shift)
{
    return finish(new clFFT_impl(fftSize, dir, window, idataType, pick_device(plat, sel, pid, did), num_streams, shift),
                  setDebug);
}
clFilter::sptr clFilter::make(int plat, int sel, int pid, int did, int decimation, const std::vector<float> &taps,
                              int nthreads, int setDebug, bool use_time)
{
    return finish(new clFilter_impl(pick_device(plat, sel, pid, did), decimation, taps, nthreads, use_time), setDebug);
}
clPolyphaseChannelizer::sptr clPolyphaseChannelizer::make(int plat, int sel, int pid, int did,
                                                          const std::vector<float> &taps, int buf_items,
                                                          int num_channels, int ninputs_per_iter,
                                                          const std::vector<int> &ch_map, int setDebug)
{
    return finish(new clPolyphaseChannelizer_impl(pick_device(plat, sel, pid, did), taps, buf_items, num_channels,
                                                  ninputs_per_iter, ch_map),
                  setDebug);
}
clXEngine::sptr clXEngine::make(int plat, int sel, int pid, int did, bool setDebug, int data_type, int polarization,
                                int num_inputs, int, int first_channel, int num_channels, int integration,
                                std::vector<std::string> antenna_list, bool output_file, std::string file_base,
                                int rollover_size_mb, bool internal_synchronizer, long sync_timestamp, std::string object_name,
                                double starting_chan_center_freq, double channel_width, bool disable_output,
                                int pipeline_integration)
{
    return finish(new clXEngine_impl(pick_device(plat, sel, pid, did), data_type, polarization, num_inputs, num_channels,
                                     integration, disable_output, pipeline_integration, output_file, file_base,
                                     rollover_size_mb, antenna_list, sync_timestamp, object_name, first_channel,
                                     starting_chan_center_freq, channel_width, internal_synchronizer),
                  setDebug ? 1 : 0);
}

clXCorrelate::sptr clXCorrelate::make(int plat, int sel, int pid, int did, bool setDebug, int num_inputs, int signal_length,
                                      int data_type, int data_size, int max_search_index, int decim_frames, bool)
{
    return finish(new clXCorrelate_impl(pick_device(plat, sel, pid, did), num_inputs, signal_length, data_type, data_size,
                                        max_search_index, decim_frames),
                  setDebug ? 1 : 0);
}
clxcorrelate_fft_vcf::sptr clxcorrelate_fft_vcf::make(int fftSize, int num_inputs, int plat, int sel, int pid, int did,
                                                      int input_type)
{
    return finish(new clxcorrelate_fft_vcf_impl(pick_device(plat, sel, pid, did), fftSize, num_inputs, input_type), 0);
}
clComplexFilter::sptr clComplexFilter::make(int plat, int sel, int pid, int did, int decimation,
                                            const std::vector<gr_complex> &taps, int, int setDebug)
{
    return finish(new clComplexFilter_impl(pick_device(plat, sel, pid, did), decimation, taps), setDebug);
}
clQuadratureDemod::sptr clQuadratureDemod::make(float gain, int plat, int sel, int pid, int did, int setDebug)
{
    return finish(new clQuadratureDemod_impl(pick_device(plat, sel, pid, did), gain), setDebug);
}
clSignalSource::sptr clSignalSource::make(int idataType, int plat, int sel, int pid, int did, double samp_rate,
                                          int waveform, double freq, float amplitude, int setDebug)
{
    return finish(new clSignalSource_impl(pick_device(plat, sel, pid, did), idataType, samp_rate, waveform, freq, amplitude),
                  setDebug);
}

} // namespace clenabled
} // namespace gr
