// Minimal stand-in for the GNU Radio runtime API that the hot-path blocks touch
// (last row): gr::block / sync_block / sync_decimator, io_signature,
// the scheduler-hint setters, message ports and the item typedefs.  It exists only so
// that the block layer compiles and its work() functions can be driven by tests in an
// image without GNU Radio; with real GNU Radio on the include path this directory is
// simply not used (same class and method names).
#ifndef CLB200_GR_STUB_BLOCK_H
#define CLB200_GR_STUB_BLOCK_H
#include <complex>
#include <cstdio>
#include <functional>
#include <cstdint>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include <pmt/pmt.h>

typedef std::complex<float> gr_complex;
typedef std::vector<int> gr_vector_int;
typedef std::vector<const void *> gr_vector_const_void_star;
typedef std::vector<void *> gr_vector_void_star;

namespace gr {

namespace thread {
typedef std::mutex mutex;
typedef std::unique_lock<std::mutex> scoped_lock;
} // namespace thread

class io_signature
{
public:
    typedef std::shared_ptr<io_signature> sptr;
    static sptr make(int min_streams, int max_streams, int sizeof_stream_item)
    {
        return sptr(new io_signature(min_streams, max_streams, sizeof_stream_item));
    }
    int min_streams() const { return d_min; }
    int max_streams() const { return d_max; }
    int sizeof_stream_item(int) const { return d_size; }

private:
    io_signature(int mn, int mx, int sz) : d_min(mn), d_max(mx), d_size(sz) {}
    int d_min, d_max, d_size;
};

// gr::logger + GR_LOG_* (gnuradio/logger.h): records the lines for tests and echoes them on stderr
struct logger {
    std::string name;
    std::vector<std::string> lines;
    void log(const char *level, const std::string &msg)
    {
        lines.push_back(std::string(level) + ": " + msg);
        fprintf(stderr, "%s :%s: %s\n", name.c_str(), level, msg.c_str());
    }
};
typedef std::shared_ptr<logger> logger_ptr;
#define GR_LOG_INFO(lg_, msg) (lg_)->log("info", (msg))
#define GR_LOG_WARN(lg_, msg) (lg_)->log("warning", (msg))
#define GR_LOG_ERROR(lg_, msg) (lg_)->log("error", (msg))

// stream tag (gnuradio/tags.h): absolute offset, key, value
struct tag_t {
    uint64_t offset = 0;
    pmt::pmt_t key, value;
};

class block
{
public:
    enum { WORK_CALLED_PRODUCE = -2, WORK_DONE = -1 };
    enum tag_propagation_policy_t { TPP_DONT = 0, TPP_ALL_TO_ALL = 1, TPP_ONE_TO_ONE = 2 };
    virtual ~block() {}
    const std::string &name() const { return d_name; }
    io_signature::sptr input_signature() const { return d_in; }
    io_signature::sptr output_signature() const { return d_out; }

    unsigned history() const { return d_history; }
    void set_history(unsigned h) { d_history = h; }
    int output_multiple() const { return d_output_multiple; }
    void set_output_multiple(int m) { d_output_multiple = m; }
    void set_alignment(int) {}
    int max_noutput_items() const { return d_max_noutput; }
    void set_max_noutput_items(int m) { d_max_noutput = m; }
    void set_tag_propagation_policy(tag_propagation_policy_t) {}
    double relative_rate() const { return d_rate; }
    void set_relative_rate(double r) { d_rate = r; }

    virtual void forecast(int noutput_items, gr_vector_int &ninput_items_required)
    {
        for (auto &n : ninput_items_required) n = noutput_items + (int)history() - 1;
    }
    virtual int general_work(int noutput_items, gr_vector_int &ninput_items,
                             gr_vector_const_void_star &input_items, gr_vector_void_star &output_items) = 0;
    virtual bool start() { return true; }
    virtual bool stop() { return true; }

    void consume(int port, int n)
    {
        if ((int)d_consumed.size() <= port) d_consumed.resize(port + 1, 0);
        d_consumed[port] += n;
    }
    void consume_each(int n) { d_consumed_each += n; }
    // test access (the real scheduler reads these through its own bookkeeping)
    int consumed_each() const { return d_consumed_each; }
    const std::vector<int> &consumed() const { return d_consumed; }

    // tags of input `port` whose offset RELATIVE to the current window lies in [start, end)
    // (block::get_tags_in_window, used by the X-engine's ATA synchroniser, lib/clXEngine_impl.cc:1170-1172)
    void get_tags_in_window(std::vector<tag_t> &v, unsigned port, uint64_t start, uint64_t end)
    {
        v.clear();
        if (port >= d_tags.size()) return;
        const uint64_t base = port < d_read.size() ? d_read[port] : 0;
        for (auto &t : d_tags[port])
            if (t.offset >= base + start && t.offset < base + end) v.push_back(t);
    }
    // test access: what the upstream block / the scheduler would do
    void test_add_tag(unsigned port, uint64_t abs_offset, pmt::pmt_t key, pmt::pmt_t value)
    {
        if (d_tags.size() <= port) d_tags.resize(port + 1);
        tag_t t;
        t.offset = abs_offset;
        t.key = key;
        t.value = value;
        d_tags[port].push_back(t);
    }
    void test_set_read_offset(unsigned port, uint64_t abs_offset)
    {
        if (d_read.size() <= port) d_read.resize(port + 1, 0);
        d_read[port] = abs_offset;
    }

    void message_port_register_out(pmt::pmt_t id) { d_ports[pmt::symbol_to_string(id)]; }
    void message_port_pub(pmt::pmt_t id, pmt::pmt_t msg) { d_ports[pmt::symbol_to_string(id)].push_back(msg); }
    std::vector<pmt::pmt_t> &published(const std::string &port) { return d_ports[port]; }

protected:
    block() : d_logger(new logger) {}          // allows pure-virtual interface sub-classes (as in GNU Radio)
    block(const std::string &name, io_signature::sptr in, io_signature::sptr out)
        : d_logger(new logger), d_name(name), d_in(in), d_out(out)
    {
        d_logger->name = name;
    }
    thread::mutex d_setlock;
    logger_ptr d_logger;

public:
    const std::vector<std::string> &test_log_lines() const { return d_logger->lines; }

private:
    std::string d_name;
    io_signature::sptr d_in, d_out;
    unsigned d_history = 1;
    int d_output_multiple = 1, d_max_noutput = 0, d_consumed_each = 0;
    double d_rate = 1.0;
    std::vector<int> d_consumed;
    std::vector<std::vector<tag_t>> d_tags;
    std::vector<uint64_t> d_read;
    std::map<std::string, std::vector<pmt::pmt_t>> d_ports;
};

class sync_block : public block
{
public:
    virtual int work(int noutput_items, gr_vector_const_void_star &input_items,
                     gr_vector_void_star &output_items) = 0;
    int general_work(int noutput_items, gr_vector_int &, gr_vector_const_void_star &in,
                     gr_vector_void_star &out) override
    {
        int r = work(noutput_items, in, out);
        if (r > 0) consume_each(r);
        return r;
    }

protected:
    sync_block() {}
    sync_block(const std::string &name, io_signature::sptr in, io_signature::sptr out) : block(name, in, out) {}
};

class sync_decimator : public sync_block
{
public:
    unsigned decimation() const { return d_decimation; }
    void set_decimation(unsigned d) { d_decimation = d; set_relative_rate(1.0 / d); }

protected:
    sync_decimator() : d_decimation(1) {}
    sync_decimator(const std::string &name, io_signature::sptr in, io_signature::sptr out, unsigned decimation)
        : sync_block(name, in, out), d_decimation(decimation)
    {
        set_relative_rate(1.0 / decimation);
    }

private:
    unsigned d_decimation;
};

} // namespace gr

namespace gnuradio {
template <class T>
std::shared_ptr<T> get_initial_sptr(T *p)
{
    return std::shared_ptr<T>(p);
}
} // namespace gnuradio
#endif
