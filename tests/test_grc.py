"""GRC block definitions (gr_clenabled_b200/grc/gen_grc.py): saved flowgraphs address a block by GRC id and
its settings by parameter id, so those are the reference's (grc/clenabled_*.block.yml; SURVEY 8b "GRC ids ->
make (must keep)").  The expected ids below were read off the reference's yml files."""
import inspect
import re

import yaml

from gr_clenabled_b200 import blocks
from gr_clenabled_b200.grc import gen_grc

DEV = ["openCLPlatform", "devices", "platformId", "deviceId"]
REFERENCE_IDS = {
    "clenabled_clMultConst": ["type", "const", "setDebug"],
    "clenabled_clAddConst": ["type", "const", "setDebug"],
    "clenabled_clComplexConjugate": ["setDebug"],
    "clenabled_clMultiply": ["type", "setDebug"],
    "clenabled_clAdd": ["type", "setDebug"],
    "clenabled_clSubtract": ["type", "setDebug"],
    "clenabled_clMultiplyConjugate": ["setDebug"],
    "clenabled_clLog10": ["setDebug", "n_val", "k_val"],
    "clenabled_clsnr": ["setDebug", "n_val", "k_val"],
    "clenabled_complextomag": ["setDebug"],
    "clenabled_complextoarg": ["setDebug"],
    "clenabled_complextomagphase": ["setDebug"],
    "clenabled_magphasetocomplex": ["setDebug"],
    "clenabled_clFFT": ["type", "fft_dir", "fft_size", "window", "shift", "num_streams", "setDebug"],
    "clenabled_clLowPassFilter": ["use_time", "decimation", "gain", "samp_rate", "cutoff_freq", "width", "win", "beta", "setDebug"],
    "clenabled_clHighPassFilter": ["use_time", "decimation", "gain", "samp_rate", "cutoff_freq", "width", "win", "beta", "setDebug"],
    "clenabled_clBandPassFilter": ["use_time", "decimation", "gain", "samp_rate", "low_cutoff_freq", "high_cutoff_freq", "width", "win", "beta", "setDebug"],
    "clenabled_clBandRejectFilter": ["use_time", "decimation", "gain", "samp_rate", "low_cutoff_freq", "high_cutoff_freq", "width", "win", "beta", "setDebug"],
    "clenabled_clRootRaisedCosineFilter": ["use_time", "samp_rate", "sym_rate", "alpha", "ntaps", "setDebug"],
    "clenabled_cltapfirfilter": ["taps", "use_time", "decimation", "samp_rate", "setDebug"],
    "clenabled_clcomplexfilter": ["taps", "decimation", "samp_rate", "setDebug"],
    "clenabled_clPolyphaseChannelizer": ["taps", "buf_items", "num_channels", "ninputs_per_iter", "chmap", "setDebug"],
    "clenabled_clQuadratureDemod": ["setDebug", "gain"],
    "clenabled_clSignalSource": ["type", "waveform", "samp_rate", "freq", "amplitude", "setDebug"],
    "clenabled_XCorrelate": ["type", "signal_length", "max_search_offset", "async", "decim_frames", "num_inputs", "setDebug"],
    "clenabled_clxcorrelate_fft_vcf": ["input_type", "vec_len", "num_inputs"],
    "clenabled_clXEngine": ["type", "sync_timestamp", "first_channel", "starting_chan_center_freq", "num_channels",
                            "channel_width", "num_inputs", "polarization", "integration", "pipeline_integration",
                            "output_file", "file_base", "rollover_size_mb", "internal_synchronizer", "object_name",
                            "antenna_list", "disable_output", "setDebug"],
}


def top_level_args(call):
    """number of top-level arguments of `cls(...)`"""
    inner = call[call.index("(") + 1:call.rindex(")")]
    depth, n = 0, 1
    for ch in inner:
        depth += ch in "([{"
        depth -= ch in ")]}"
        n += (ch == "," and depth == 0)
    return n


def test_every_reference_grc_id_is_generated_with_its_parameter_ids(tmp_path):
    files = gen_grc.main(str(tmp_path))
    defs = {}
    for f in files:
        d = yaml.safe_load(open(f))
        defs[d["id"]] = d
    assert set(defs) == set(REFERENCE_IDS)
    for gid, want in REFERENCE_IDS.items():
        d = defs[gid]
        have = [p["id"] for p in d["parameters"]]
        assert set(DEV + want) <= set(have), (gid, set(DEV + want) - set(have))
        assert len(have) == len(set(have))
        make = d["templates"]["make"].strip()
        m = re.match(r"clenabled\.(\w+)\(", make)
        assert m, make
        cls = getattr(blocks, m.group(1))                   # the python name the flowgraph instantiates exists
        sig = inspect.signature(cls.__init__)
        npos = len([p for p in sig.parameters.values() if p.name != "self"])
        nreq = len([p for p in sig.parameters.values() if p.name != "self" and p.default is inspect.Parameter.empty])
        assert nreq <= top_level_args(make) <= npos, (gid, make)
        for ref in re.findall(r"\$\{\s*(\w+)", make):      # every ${param} used by make() is a declared parameter
            assert ref in have, (gid, ref)
        assert d["file_format"] == 1


def test_import_clenabled_serves_every_class_the_grc_templates_name(tmp_path):
    """GRC-generated flowgraphs do `import clenabled` and call `clenabled.<Class>(...)`"""
    import clenabled
    for f in gen_grc.main(str(tmp_path)):
        make = yaml.safe_load(open(f))["templates"]["make"].strip()
        cls = re.match(r"clenabled\.(\w+)\(", make).group(1)
        assert getattr(clenabled, cls) is getattr(blocks, cls)
