"""Regenerates tests/golden/ref_constants.json from the REFERENCE's headers and GRC files
(needs /root/reference; run in the build container):

    python tests/golden/make_constants.py

The numeric ids cross the drop-in boundary as plain ints (saved flowgraphs, reference-API callers), so
tests/test_constants.py pins ours to these.
"""
import json
import os
import re

import yaml

REF = os.environ.get("CLB200_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def defines(path, prefixes):
    out = {}
    for line in open(path):
        m = re.match(r"\s*#define\s+(\w+)\s+(-?\d+)\s*$", line)
        if m and m.group(1).startswith(prefixes):
            out[m.group(1)] = int(m.group(2))
    return out


def enum_values(yml, param_id, attr):
    with open(yml) as f:
        d = yaml.safe_load(f)
    for p in d["parameters"]:
        if p["id"] == param_id:
            return [int(v) for v in p["option_attributes"][attr]]
    raise KeyError((yml, param_id))


PYBIND_CLASSES = ["clMathConst", "clMathOp", "clLog", "clSNR", "clComplexToMag", "clComplexToArg", "clComplexToMagPhase",
                  "clMagPhaseToComplex", "clFFT", "clFilter", "clPolyphaseChannelizer", "clXEngine", "clXCorrelate",
                  "clxcorrelate_fft_vcf", "clComplexFilter", "clQuadratureDemod", "clSignalSource"]


def pybind_init_args(cls):
    """keyword names of the constructor as python/bindings/<cls>_python.cc binds them, in order"""
    text = open(os.path.join(REF, "python", "bindings", cls + "_python.cc")).read()
    init = text[text.index("py::init("):]
    init = init[:init.index("D(" + cls + ",make)")] if ("D(" + cls + ",make)") in init else init[:init.index(")\n\n")]
    return re.findall(r'py::arg\("(\w+)"\)', init)


def main():
    inc = os.path.join(REF, "include", "clenabled")
    out = {
        "GRCLBase.h": defines(os.path.join(inc, "GRCLBase.h"), ("DTYPE_", "OCLTYPE_", "OCLDEVICESELECTOR_")),
        "clMathOpTypes.h": defines(os.path.join(inc, "clMathOpTypes.h"), ("MATHOP_",)),
        "grc": {
            "clenabled_clXEngine.type.data_type": enum_values(os.path.join(REF, "grc", "clenabled_clXEngine.block.yml"), "type", "data_type"),
            "clenabled_clFFT.type.datatype": enum_values(os.path.join(REF, "grc", "clenabled_clFFT.block.yml"), "type", "datatype"),
            "clenabled_clMultConst.type.datatype": enum_values(os.path.join(REF, "grc", "clenabled_clMultConst.block.yml"), "type", "datatype"),
        },
    }
    out["pybind_module"] = "clenabled_python"
    out["pybind_init_args"] = {c: pybind_init_args(c) for c in PYBIND_CLASSES}
    with open(os.path.join(HERE, "ref_constants.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
        f.write("\n")
    print(json.dumps(out, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
