// bindings.cc -- the pybind11 module of the OOT: `clenabled_python`, same module name, class names and keyword
// names as the reference's python/bindings/*_python.cc (python_bindings.cc:57-93; the keyword lists below are the
// reference's, position by position -- including clFFT's, whose names are shifted against the C++ parameters
// (clFFT_python.cc:39-51), because existing Python flowgraphs pass them by those names).
//
// With real GNU Radio on the include path the classes derive from the gnuradio.gr bindings of gr::block /
// sync_block / sync_decimator exactly like the reference's.  Against the in-repo GNU Radio stub (this image has no
// GNU Radio) the module binds the stub's gr::block itself and adds a `general_work(noutput_items, inputs, outputs)`
// driver over numpy arrays, which is what the tests use to run the blocks from Python.
#include <pybind11/complex.h>
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <clenabled/blocks.h>

namespace py = pybind11;
using namespace gr::clenabled;

#if __has_include(<gnuradio/basic_block.h>)
#define CLB200_REAL_GNURADIO 1
#else
#define CLB200_REAL_GNURADIO 0
#endif

namespace {

#if !CLB200_REAL_GNURADIO
py::object pmt_to_python(const pmt::pmt_t &p)
{
    using K = pmt::pmt_base;
    switch (p->kind) {
    case K::SYMBOL: return py::str(p->sym);
    case K::PAIR: return py::make_tuple(pmt_to_python(p->car), pmt_to_python(p->cdr));
    case K::C32VECTOR: return py::array_t<std::complex<float>>(p->c32.size(), p->c32.data());
    case K::F32VECTOR: return py::array_t<float>(p->f32.size(), p->f32.data());
    case K::S32VECTOR: return py::array_t<int32_t>(p->s32.size(), p->s32.data());
    case K::UINT64: return py::int_(p->u64);
    case K::DICT: {
        py::dict d;
        for (auto &kv : p->dict) d[py::str(kv.first)] = pmt_to_python(kv.second);
        return std::move(d);
    }
    default: return py::none();
    }
}

// what the scheduler does for one call: raw item pointers in, items produced back
int drive(gr::block &b, int noutput_items, std::vector<py::array> inputs, std::vector<py::array> outputs)
{
    gr_vector_const_void_star in;
    gr_vector_void_star out;
    gr_vector_int ninput;
    for (auto &a : inputs) {
        if (!(a.flags() & py::array::c_style)) throw std::invalid_argument("inputs must be C-contiguous arrays");
        in.push_back(a.data());
        const size_t isz = (size_t)b.input_signature()->sizeof_stream_item((int)in.size() - 1);
        ninput.push_back((int)(isz ? (size_t)a.nbytes() / isz : 0));
    }
    for (auto &a : outputs) {
        if (!(a.flags() & py::array::c_style) || !a.writeable()) throw std::invalid_argument("outputs must be writeable C-contiguous arrays");
        out.push_back(a.mutable_data());
    }
    py::gil_scoped_release nogil;
    return b.general_work(noutput_items, ninput, in, out);
}
#endif

void bind_bases(py::module &m)
{
#if CLB200_REAL_GNURADIO
    py::module::import("gnuradio.gr");                       // gr::basic_block / block / sync_block / sync_decimator
#else
    py::class_<gr::block, std::shared_ptr<gr::block>>(m, "block")
        .def("name", &gr::block::name)
        .def("history", &gr::block::history)
        .def("output_multiple", &gr::block::output_multiple)
        .def("relative_rate", &gr::block::relative_rate)
        .def("start", &gr::block::start)
        .def("stop", &gr::block::stop)
        .def("general_work", &drive, py::arg("noutput_items"), py::arg("inputs"), py::arg("outputs"),
             "one scheduler call: inputs / outputs are numpy arrays over the stream items; returns the items produced")
        .def("consumed_each", &gr::block::consumed_each)
        .def("consumed", &gr::block::consumed)
        .def("log_lines", &gr::block::test_log_lines)
        .def("published", [](gr::block &b, const std::string &port) {
            py::list l;
            for (auto &msg : b.published(port)) l.append(pmt_to_python(msg));
            return l;
        }, py::arg("port"), "messages published on a message port so far");
    py::class_<gr::sync_block, gr::block, std::shared_ptr<gr::sync_block>>(m, "sync_block");
    py::class_<gr::sync_decimator, gr::sync_block, std::shared_ptr<gr::sync_decimator>>(m, "sync_decimator")
        .def("decimation", &gr::sync_decimator::decimation);
#endif
}

#if CLB200_REAL_GNURADIO
#define SYNC_BASES gr::sync_block, gr::block, gr::basic_block
#define DECIM_BASES gr::sync_decimator
#define BLOCK_BASES gr::block, gr::basic_block
#else
#define SYNC_BASES gr::sync_block
#define DECIM_BASES gr::sync_decimator
#define BLOCK_BASES gr::block
#endif

} // namespace

PYBIND11_MODULE(clenabled_python, m)
{
    m.doc() = "gr-clenabled blocks on CUDA (sm_100a): the reference's pybind11 API over libclenabled_b200";
    bind_bases(m);

    py::class_<clMathConst, SYNC_BASES, std::shared_ptr<clMathConst>>(m, "clMathConst")
        .def(py::init(&clMathConst::make), py::arg("idataType"), py::arg("openCLPlatformType"), py::arg("devSelector"),
             py::arg("platformId"), py::arg("devId"), py::arg("fValue"), py::arg("operatorType"), py::arg("setDebug") = 0)
        .def("k", &clMathConst::k)
        .def("set_k", &clMathConst::set_k, py::arg("newValue"));
    py::class_<clMathOp, SYNC_BASES, std::shared_ptr<clMathOp>>(m, "clMathOp")
        .def(py::init(&clMathOp::make), py::arg("idataType"), py::arg("openCLPlatformType"), py::arg("devSelector"),
             py::arg("platformId"), py::arg("devId"), py::arg("operatorType"), py::arg("setDebug") = 0);
    py::class_<clLog, SYNC_BASES, std::shared_ptr<clLog>>(m, "clLog")
        .def(py::init(&clLog::make), py::arg("openCLPlatformType"), py::arg("devSelector"), py::arg("platformId"),
             py::arg("devId"), py::arg("nValue"), py::arg("kValue"), py::arg("setDebug") = 0);
    py::class_<clSNR, SYNC_BASES, std::shared_ptr<clSNR>>(m, "clSNR")
        .def(py::init(&clSNR::make), py::arg("openCLPlatformType"), py::arg("devSelector"), py::arg("platformId"),
             py::arg("devId"), py::arg("nValue"), py::arg("kValue"), py::arg("setDebug") = 0);
    py::class_<clComplexToMag, SYNC_BASES, std::shared_ptr<clComplexToMag>>(m, "clComplexToMag")
        .def(py::init(&clComplexToMag::make), py::arg("openCLPlatformType"), py::arg("devSelector"), py::arg("platformId"),
             py::arg("devId"), py::arg("setDebug") = 0);
    py::class_<clComplexToArg, SYNC_BASES, std::shared_ptr<clComplexToArg>>(m, "clComplexToArg")
        .def(py::init(&clComplexToArg::make), py::arg("openCLPlatformType"), py::arg("devSelector"), py::arg("platformId"),
             py::arg("devId"), py::arg("setDebug") = 0);
    py::class_<clComplexToMagPhase, SYNC_BASES, std::shared_ptr<clComplexToMagPhase>>(m, "clComplexToMagPhase")
        .def(py::init(&clComplexToMagPhase::make), py::arg("openCLPlatformType"), py::arg("devSelector"),
             py::arg("platformId"), py::arg("devId"), py::arg("setDebug") = 0);
    py::class_<clMagPhaseToComplex, SYNC_BASES, std::shared_ptr<clMagPhaseToComplex>>(m, "clMagPhaseToComplex")
        .def(py::init(&clMagPhaseToComplex::make), py::arg("openCLPlatformType"), py::arg("devSelector"),
             py::arg("platformId"), py::arg("devId"), py::arg("setDebug") = 0);
    // keyword names as the reference binds them (clFFT_python.cc:39-51): positions 5..8 of make() are
    // (openCLPlatformType, devSelector, platformId, devId) but are NAMED (devSelector, platformId, devId, openCLPlatformType=4)
    py::class_<clFFT, SYNC_BASES, std::shared_ptr<clFFT>>(m, "clFFT")
        .def(py::init(&clFFT::make), py::arg("fftSize"), py::arg("clFFTDir"), py::arg("window"), py::arg("idataType"),
             py::arg("devSelector"), py::arg("platformId"), py::arg("devId"), py::arg("openCLPlatformType") = 4,
             py::arg("setDebug") = 0, py::arg("num_streams") = 1, py::arg("shift") = false);
    py::class_<clFilter, DECIM_BASES, std::shared_ptr<clFilter>>(m, "clFilter")
        .def(py::init(&clFilter::make), py::arg("openclPlatform"), py::arg("devSelector"), py::arg("platformId"),
             py::arg("devId"), py::arg("decimation"), py::arg("taps"), py::arg("nthreads") = 1, py::arg("setDebug") = 0,
             py::arg("use_time") = false)
        .def("set_taps2", &clFilter::set_taps2, py::arg("taps"))
        .def("taps", &clFilter::taps)
        .def("set_nthreads", &clFilter::set_nthreads, py::arg("n"))
        .def("nthreads", &clFilter::nthreads);
    py::class_<clPolyphaseChannelizer, BLOCK_BASES, std::shared_ptr<clPolyphaseChannelizer>>(m, "clPolyphaseChannelizer")
        .def(py::init(&clPolyphaseChannelizer::make), py::arg("openCLPlatformType"), py::arg("devSelector"),
             py::arg("platformId"), py::arg("devId"), py::arg("taps"), py::arg("buf_items"), py::arg("num_channels"),
             py::arg("ninputs_per_iter"), py::arg("ch_map"), py::arg("setDebug") = 0);
    py::class_<clXEngine, BLOCK_BASES, std::shared_ptr<clXEngine>>(m, "clXEngine")
        .def(py::init(&clXEngine::make), py::arg("openCLPlatformType"), py::arg("devSelector"), py::arg("platformId"),
             py::arg("devId"), py::arg("setDebug"), py::arg("data_type"), py::arg("polarization"), py::arg("num_inputs"),
             py::arg("output_format"), py::arg("first_channel"), py::arg("num_channels"), py::arg("integration"),
             py::arg("antenna_list"), py::arg("output_file") = false, py::arg("file_base") = "",
             py::arg("rollover_size_mb") = 0, py::arg("internal_synchronizer") = false, py::arg("sync_timestamp") = 0,
             py::arg("object_name") = "", py::arg("starting_chan_center_freq") = 0., py::arg("channel_width") = 0.,
             py::arg("disable_output") = false, py::arg("pipeline_integration") = 0);
    py::class_<clXCorrelate, SYNC_BASES, std::shared_ptr<clXCorrelate>>(m, "clXCorrelate")
        .def(py::init(&clXCorrelate::make), py::arg("openCLPlatformType"), py::arg("devSelector"), py::arg("platformId"),
             py::arg("devId"), py::arg("setDebug"), py::arg("num_inputs"), py::arg("signal_length"), py::arg("data_type"),
             py::arg("data_size"), py::arg("max_search_index"), py::arg("decim_frames"), py::arg("async") = false);
    py::class_<clxcorrelate_fft_vcf, SYNC_BASES, std::shared_ptr<clxcorrelate_fft_vcf>>(m, "clxcorrelate_fft_vcf")
        .def(py::init(&clxcorrelate_fft_vcf::make), py::arg("fftSize"), py::arg("num_inputs"), py::arg("openCLPlatformType"),
             py::arg("devSelector"), py::arg("platformId"), py::arg("devId"), py::arg("input_type") = 1);
    py::class_<clComplexFilter, DECIM_BASES, std::shared_ptr<clComplexFilter>>(m, "clComplexFilter")
        .def(py::init(&clComplexFilter::make), py::arg("openclPlatform"), py::arg("devSelector"), py::arg("platformId"),
             py::arg("devId"), py::arg("decimation"), py::arg("taps"), py::arg("nthreads") = 1, py::arg("setDebug") = 0)
        .def("set_taps2", &clComplexFilter::set_taps2, py::arg("taps"))
        .def("taps", &clComplexFilter::taps);
    py::class_<clQuadratureDemod, SYNC_BASES, std::shared_ptr<clQuadratureDemod>>(m, "clQuadratureDemod")
        .def(py::init(&clQuadratureDemod::make), py::arg("gain"), py::arg("openCLPlatformType"), py::arg("devSelector"),
             py::arg("platformId"), py::arg("devId"), py::arg("setDebug") = 0);
    py::class_<clSignalSource, SYNC_BASES, std::shared_ptr<clSignalSource>>(m, "clSignalSource")
        .def(py::init(&clSignalSource::make), py::arg("idataType"), py::arg("openCLPlatformType"), py::arg("devSelector"),
             py::arg("platformId"), py::arg("devId"), py::arg("samp_rate"), py::arg("waveform"), py::arg("freq"),
             py::arg("amplitude"), py::arg("setDebug") = 0);
}
