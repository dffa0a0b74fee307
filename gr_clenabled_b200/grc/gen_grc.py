#!/usr/bin/env python
"""Generates the GNU Radio Companion block definitions (*.block.yml) of the B200 build.

Saved flowgraphs refer to a block by its GRC id and to its settings by parameter id, so both are
the reference's (grc/clenabled_*.block.yml: ids listed in SURVEY 8b): an existing .grc file opens
unchanged and instantiates the same python names (`clenabled.clFFT(...)`) with the same argument
lists.  Everything else is produced from the table below -- one row per block -- instead of 27
hand-kept files; labels say CUDA.  One make() form per block: the reference's separate
"Any device" branch hard-codes (1,0,0) and, for the Const blocks, drops the constant
(grc/clenabled_clMultConst.block.yml:61) -- not reproduced.

usage: python gen_grc.py [outdir]      (default: gr_clenabled_b200/lib/grc)
"""
import os
import sys

DEV = "${openCLPlatform},${devices},${platformId},${deviceId}"


def enum(pid, label, options, labels=None, **kw):
    d = {"id": pid, "label": label, "dtype": "enum", "options": [str(o) for o in options]}
    if labels:
        d["option_labels"] = list(labels)
    d.update(kw)
    return d


def par(pid, label, dtype, default=None, **kw):
    d = {"id": pid, "label": label, "dtype": dtype}
    if default is not None:
        d["default"] = str(default)
    d.update(kw)
    return d


def device_params():
    """the four device-selection fields every block carries (GRCLBase.h:64-70); CUDA: one platform"""
    return [
        enum("openCLPlatform", "Device Type", [1, 2, 3, 4], ["GPU", "Accelerator", "CPU (unsupported)", "Any"]),
        enum("devices", "CUDA Device", [1, 2], ["First", "Specific"], option_attributes={"hide_specific": ["all", ""]}),
        enum("platformId", "Platform Id", [0, 1, 2, 3], hide="${ devices.hide_specific }"),
        enum("deviceId", "Device Id (CUDA ordinal)", [0, 1, 2, 3, 4, 5, 6, 7], hide="${ devices.hide_specific }"),
    ]


DEBUG = enum("setDebug", "Debug", [0, 1], ["Off", "On"])
TYPE3 = enum("type", "Data Type", ["complex", "float", "int"], hide="part",
             option_attributes={"datatype": ["1", "2", "3"], "input": ["complex", "float", "int"],
                                "output": ["complex", "float", "int"]})
USE_TIME = enum("use_time", "Filter Domain", ["False", "True"], ["Frequency (FFT)", "Time (FIR)"])


def stream(dtype, vlen=None, mult=None, **kw):
    d = {"domain": "stream", "dtype": dtype}
    if vlen:
        d["vlen"] = vlen
    if mult:
        d["multiplicity"] = mult
    d.update(kw)
    return d


def msg(pid):
    return {"domain": "message", "id": pid, "optional": True}


def filt(gid, label, design, extra):
    """the firdes wrappers: all of them are clFilter with designed taps"""
    return dict(id=gid, label=label, dev_first=True,
                params=[USE_TIME, par("decimation", "Decimation", "int", 1), par("gain", "Gain", "real", 1),
                        par("samp_rate", "Sample Rate", "real", "samp_rate")] + extra + [DEBUG],
                ins=[stream("complex")], outs=[stream("complex")],
                imports="import clenabled\nfrom gnuradio.filter import firdes\nfrom gnuradio.fft import window",
                make="clenabled.clFilter(%s,${decimation},%s,1,${setDebug},${use_time})" % (DEV, design),
                callbacks=["set_taps2(%s)" % design])


WIN = [par("win", "Window", "raw", "window.WIN_HAMMING"), par("beta", "Beta", "real", 6.76)]
T1 = [par("cutoff_freq", "Cutoff Freq", "real"), par("width", "Transition Width", "real")] + WIN
T2 = [par("low_cutoff_freq", "Low Cutoff Freq", "real"), par("high_cutoff_freq", "High Cutoff Freq", "real"),
      par("width", "Transition Width", "real")] + WIN


def mathconst(gid, label, op, typed=True, const=True):
    t = "${type.datatype}" if typed else "1"
    k = "${const}" if const else "0"
    io = "${ type.input }" if typed else "complex"
    return dict(id=gid, label=label, type_first=typed,
                params=([par("const", "Constant", "${ type.input }" if typed else "complex", 1)] if const else []) + [DEBUG],
                ins=[stream(io)], outs=[stream("${ type.output }" if typed else "complex")],
                make="clenabled.clMathConst(%s,%s,%s,%d,${setDebug})" % (t, DEV, k, op),
                callbacks=["set_k(${const})"] if const else None)


def mathop(gid, label, op, typed=True):
    t = "${type.datatype}" if typed else "1"
    io = "${ type.input }" if typed else "complex"
    return dict(id=gid, label=label, type_first=typed, params=[DEBUG], ins=[stream(io), stream(io)],
                outs=[stream("${ type.output }" if typed else "complex")],
                make="clenabled.clMathOp(%s,%s,%d,${setDebug})" % (t, DEV, op))


def simple(gid, label, cls, ins, outs, extra=None, args=""):
    return dict(id=gid, label=label, params=[DEBUG] + (extra or []), ins=ins, outs=outs,
                make="clenabled.%s(%s%s,${setDebug})" % (cls, DEV, args))


NK = [par("n_val", "n", "float", 1), par("k_val", "k", "float", 0)]
XE_TYPE = enum("type", "Input Type", ["complex", "ichar", "packed4"], ["Complex", "Byte (IChar)", "Packed 4-bit XY"],
               option_attributes={"data_type": ["1", "5", "6"], "input_format": ["complex", "byte", "byte"]})
XC_TYPE = enum("type", "Input Type", ["complex", "float"],
               option_attributes={"data_type": ["1", "2"], "size": ["8", "4"]})
BOOL = lambda pid, label: enum(pid, label, ["False", "True"], ["No", "Yes"])     # noqa: E731

BLOCKS = [
    mathconst("clenabled_clMultConst", "CUDA Multiply Const", 1),
    mathconst("clenabled_clAddConst", "CUDA Add Const", 2),
    mathconst("clenabled_clComplexConjugate", "CUDA Complex Conjugate", 4, typed=False, const=False),
    mathop("clenabled_clMultiply", "CUDA Multiply", 1),
    mathop("clenabled_clAdd", "CUDA Add", 2),
    mathop("clenabled_clSubtract", "CUDA Subtract", 3),
    mathop("clenabled_clMultiplyConjugate", "CUDA Multiply Conjugate", 5, typed=False),
    simple("clenabled_clLog10", "CUDA Log10", "clLog", [stream("float")], [stream("float")], NK, ",${n_val},${k_val}"),
    simple("clenabled_clsnr", "CUDA SNR Helper", "clSNR", [stream("float"), stream("float")], [stream("float")], NK,
           ",${n_val},${k_val}"),
    simple("clenabled_complextomag", "CUDA Complex To Mag", "clComplexToMag", [stream("complex")], [stream("float")]),
    simple("clenabled_complextoarg", "CUDA Complex To Arg", "clComplexToArg", [stream("complex")], [stream("float")]),
    simple("clenabled_complextomagphase", "CUDA Complex To Mag Phase", "clComplexToMagPhase", [stream("complex")],
           [stream("float"), stream("float")]),
    simple("clenabled_magphasetocomplex", "CUDA Mag Phase To Complex", "clMagPhaseToComplex",
           [stream("float"), stream("float")], [stream("complex")]),
    dict(id="clenabled_clFFT", label="CUDA FFT", type_first=True,
         type_param=enum("type", "Input Type", ["complex", "float"], hide="part",
                         option_attributes={"datatype": ["1", "2"], "input": ["complex", "float"]}),
         params=[enum("fft_dir", "Direction", [-1, 1], ["Forward", "Reverse"]), par("fft_size", "FFT Size", "int", 1024),
                 par("window", "Window", "real_vector", "window.blackmanharris(1024)"), BOOL("shift", "Shift"),
                 par("num_streams", "Num Streams", "int", 1), DEBUG],
         ins=[stream("${ type.input }", "${ fft_size }", "${num_streams}")],
         outs=[stream("complex", "${ fft_size }", "${num_streams}")],
         imports="from gnuradio.fft import window\nimport clenabled",
         make="clenabled.clFFT(${fft_size},${fft_dir},${window},${type.datatype},%s,${setDebug},${num_streams},${shift})" % DEV),
    filt("clenabled_clLowPassFilter", "CUDA Low Pass Filter",
         "firdes.low_pass(${gain}, ${samp_rate}, ${cutoff_freq}, ${width}, ${win}, ${beta})", T1),
    filt("clenabled_clHighPassFilter", "CUDA High Pass Filter",
         "firdes.high_pass(${gain}, ${samp_rate}, ${cutoff_freq}, ${width}, ${win}, ${beta})", T1),
    filt("clenabled_clBandPassFilter", "CUDA Band Pass Filter",
         "firdes.band_pass(${gain}, ${samp_rate}, ${low_cutoff_freq}, ${high_cutoff_freq}, ${width}, ${win}, ${beta})", T2),
    filt("clenabled_clBandRejectFilter", "CUDA Band Reject Filter",
         "firdes.band_reject(${gain}, ${samp_rate}, ${low_cutoff_freq}, ${high_cutoff_freq}, ${width}, ${win}, ${beta})", T2),
    filt("clenabled_clRootRaisedCosineFilter", "CUDA Root Raised Cosine Filter",
         "firdes.root_raised_cosine(${gain}, ${samp_rate}, ${sym_rate}, ${alpha}, ${ntaps})",
         [par("sym_rate", "Symbol Rate", "real", 1.0), par("alpha", "Alpha", "real", 0.35),
          par("ntaps", "Num Taps", "int", "11*samp_rate")]),
    dict(id="clenabled_cltapfirfilter", label="CUDA Tap-based Filter", dev_first=True,
         params=[par("taps", "Taps", "real_vector"), USE_TIME, par("decimation", "Decimation", "int", 1),
                 par("samp_rate", "Sample Rate", "real", "samp_rate"), DEBUG],
         ins=[stream("complex")], outs=[stream("complex")],
         make="clenabled.clFilter(%s,${decimation},${taps},1,${setDebug},${use_time})" % DEV, callbacks=["set_taps2(${taps})"]),
    dict(id="clenabled_clcomplexfilter", label="CUDA Complex FIR Filter", dev_first=True,
         params=[par("taps", "Taps", "complex_vector"), par("decimation", "Decimation", "int", 1),
                 par("samp_rate", "Sample Rate", "real", "samp_rate"), DEBUG],
         ins=[stream("complex")], outs=[stream("complex")],
         make="clenabled.clComplexFilter(%s,${decimation},${taps},1,${setDebug})" % DEV, callbacks=["set_taps2(${taps})"]),
    dict(id="clenabled_clPolyphaseChannelizer", label="CUDA Polyphase Channelizer", dev_first=True,
         params=[par("taps", "Taps", "real_vector"), par("buf_items", "Buffer Items", "int"),
                 par("num_channels", "Channels", "int"), par("ninputs_per_iter", "Inputs Per Iteration", "int"),
                 par("chmap", "Channel Map", "int_vector"), DEBUG],
         ins=[stream("complex")], outs=[stream("complex")],
         make="clenabled.clPolyphaseChannelizer(%s, ${taps}, ${buf_items}, ${num_channels}, ${ninputs_per_iter}, ${chmap})" % DEV),
    dict(id="clenabled_clQuadratureDemod", label="CUDA Quadrature Demod", dev_first=True,
         params=[DEBUG, par("gain", "Gain", "float", 1.0)], ins=[stream("complex")], outs=[stream("float")],
         make="clenabled.clQuadratureDemod(${gain},%s,${setDebug})" % DEV),
    dict(id="clenabled_clSignalSource", label="CUDA Signal Source", type_first=True,
         type_param=enum("type", "Output Type", ["complex", "float"], hide="part",
                         option_attributes={"datatype": ["1", "2"], "output": ["complex", "float"]}),
         params=[enum("waveform", "Waveform", [1, 2], ["Cosine", "Sine"]), par("samp_rate", "Sample Rate", "float", "samp_rate"),
                 par("freq", "Frequency", "float", 1000), par("amplitude", "Amplitude", "float", 1.0), DEBUG],
         ins=[], outs=[stream("${ type.output }")],
         make="clenabled.clSignalSource(${type.datatype},%s,${samp_rate},${waveform}, ${freq}, ${amplitude},${setDebug})" % DEV),
    dict(id="clenabled_XCorrelate", label="CUDA Ref Correlate TD", dev_last=True,
         params=[XC_TYPE, par("signal_length", "Signal Length", "int", 8192), par("max_search_offset", "Max Search Offset", "int", 512),
                 BOOL("async", "Async"), par("decim_frames", "Keep 1 in N Frames", "int", 4), par("num_inputs", "Num Inputs", "int", 2)],
         tail=[DEBUG], ins=[stream("${ type }", None, "${ num_inputs }")], outs=[msg("corr")],
         make="clenabled.clXCorrelate(%s,${setDebug},${num_inputs},${signal_length},${type.data_type},${type.size},"
              "${max_search_offset},${decim_frames},${async})" % DEV),
    dict(id="clenabled_clxcorrelate_fft_vcf", label="CUDA Ref Correlate FD", dev_last=True,
         params=[enum("input_type", "Input Type", [1, 2], ["FFT", "Time Series"]), par("vec_len", "Vector Length", "int", 1024),
                 par("num_inputs", "Num Inputs", "int", 2)],
         ins=[stream("complex", "${vec_len}", "${ num_inputs }")], outs=[stream("float", "${vec_len}", "${ num_inputs - 1 }")],
         make="clenabled.clxcorrelate_fft_vcf(${vec_len},${num_inputs},%s,${input_type})" % DEV),
    dict(id="clenabled_clXEngine", label="CUDA X-Engine", dev_first=True,
         params=[XE_TYPE, par("sync_timestamp", "Sync Timestamp", "int", 0), par("first_channel", "First Channel", "int", 0),
                 par("starting_chan_center_freq", "First Channel Center Freq", "float", 0),
                 par("num_channels", "Channels", "int", 256), par("channel_width", "Channel Width", "float", 0),
                 par("num_inputs", "Num Inputs", "int", 2), enum("polarization", "Polarizations", [1, 2]),
                 par("integration", "Integration Frames", "int", 10000), par("pipeline_integration", "Pipeline Integrations", "int", 0),
                 BOOL("output_file", "Output To File"), par("file_base", "File Base", "string", ""),
                 par("rollover_size_mb", "Rollover Size (MB)", "int", 0), BOOL("internal_synchronizer", "ATA SNAP Synchronizer"),
                 par("object_name", "Object Name", "string", ""), par("antenna_list", "Antenna List", "string", ""),
                 BOOL("disable_output", "Disable Output"), DEBUG],
         ins=[stream("${ type.input_format }", "${ (num_channels if type.data_type == 1 else num_channels*2) }", "${num_inputs}"),
              stream("${ type.input_format }", "${ (num_channels if type.data_type == 1 else num_channels*2) }",
                     "${ (0 if ((polarization == '1') or (type.data_type == 6)) else num_inputs) }",
                     optional="${ (True if ((polarization == '1') or (type.data_type == 6)) else False) }")],
         outs=[msg("xcorr"), msg("sync")],
         make="clenabled.clXEngine(%s, ${setDebug}, ${type.data_type}, ${polarization}, ${num_inputs}, 1, ${first_channel}, "
              "${num_channels}, ${integration}, ${antenna_list}.replace(' ','').split(','), ${output_file},${file_base},"
              "${rollover_size_mb},${internal_synchronizer},${sync_timestamp}, ${object_name}, ${starting_chan_center_freq}, "
              "${channel_width}, ${disable_output},${pipeline_integration})" % DEV),
]


def emit_value(v, ind):
    pad = "    " * ind
    if isinstance(v, dict):
        return "\n" + "".join("%s%s:%s\n" % (pad, k, emit_value(x, ind + 1)) for k, x in v.items()).rstrip("\n")
    if isinstance(v, list) and v and all(not isinstance(x, (dict, list)) for x in v):
        return " [" + ", ".join(q(x) for x in v) + "]"
    if isinstance(v, list):
        return " []"
    return " " + q(v)


def q(x):
    s = str(x)
    plain = s and all(c.isalnum() or c in "_-. ()/" for c in s) and not s[0].isdigit() and s not in ("True", "False", "Off", "On", "No", "Yes")
    return s if plain else "'" + s.replace("'", "''") + "'"


def emit_list(items):
    out = []
    for it in items:
        first = True
        for k, v in it.items():
            out.append("%s   %s:%s" % ("-" if first else " ", k, emit_value(v, 2)))
            first = False
    return "\n".join(out)


def render(b):
    params = []
    if b.get("type_first"):
        params.append(b.get("type_param", TYPE3))
    if b.get("dev_last"):
        params += b["params"] + device_params() + b.get("tail", [])
    else:
        params += device_params() + b["params"]
    lines = ["# generated by gr_clenabled_b200/grc/gen_grc.py -- do not edit", "", "id: " + b["id"], "label: " + b["label"],
             "category: '[CUDA-Enabled (B200)]'", "", "parameters:", emit_list(params), ""]
    if b["ins"]:
        lines += ["inputs:", emit_list(b["ins"]), ""]
    if b["outs"]:
        lines += ["outputs:", emit_list(b["outs"]), ""]
    lines += ["templates:"]
    imp = b.get("imports", "import clenabled")
    if "\n" in imp:
        lines += ["    imports: |-"] + ["        " + l for l in imp.split("\n")]
    else:
        lines += ["    imports: " + imp]
    lines += ["    make: |-", "        " + b["make"]]
    if b.get("callbacks"):
        lines += ["    callbacks:"] + ["    - " + c for c in b["callbacks"]]
    lines += ["", "documentation: |-",
              "    B200 (sm_100a) CUDA implementation behind the gr-clenabled block of the same id;",
              "    devices: First = CUDA ordinal 0, Specific = the ordinal in Device Id.", "", "file_format: 1", ""]
    return "\n".join(lines)


def main(outdir=None):
    here = os.path.dirname(os.path.abspath(__file__))
    outdir = outdir or os.path.join(os.path.dirname(here), "lib", "grc")
    os.makedirs(outdir, exist_ok=True)
    names = []
    for b in BLOCKS:
        fn = os.path.join(outdir, "b200_%s.block.yml" % b["id"].replace("clenabled_", ""))
        with open(fn, "w") as f:
            f.write(render(b))
        names.append(fn)
    return names


if __name__ == "__main__":
    for n in main(sys.argv[1] if len(sys.argv) > 1 else None):
        print(n)
