"""The numeric ids that cross the drop-in boundary as plain ints (data types, operator codes, platform /
device selectors) equal the reference's (include/clenabled/GRCLBase.h:57-70, clMathOpTypes.h:11-20, the
GRC enums).  tests/golden/ref_constants.json was read off the reference tree by
tests/golden/make_constants.py; when the tree is present the fixture itself is re-checked against it."""
import json
import os
import re

import pytest
import yaml

from gr_clenabled_b200 import capi
from gr_clenabled_b200.grc import gen_grc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
with open(os.path.join(ROOT, "tests", "golden", "ref_constants.json")) as f:
    REF = json.load(f)


def _defines(path):
    out = {}
    for line in open(path):
        m = re.match(r"\s*#define\s+(\w+)\s+\(?(-?\d+)\)?", line)
        if m:
            out[m.group(1)] = int(m.group(2))
    return out


def test_block_header_constants_equal_the_reference():
    ours = _defines(os.path.join(ROOT, "gr_clenabled_b200", "host", "include", "clenabled", "blocks.h"))
    for hdr in ("GRCLBase.h", "clMathOpTypes.h"):
        for name, val in REF[hdr].items():
            if name in ours:
                assert ours[name] == val, name
    for name in ("DTYPE_COMPLEX", "DTYPE_FLOAT", "DTYPE_INT", "DTYPE_SHORT", "DTYPE_BYTE", "DTYPE_PACKEDXY",
                 "OCLTYPE_GPU", "OCLTYPE_ANY", "OCLDEVICESELECTOR_FIRST", "OCLDEVICESELECTOR_SPECIFIC",
                 "MATHOP_MULTIPLY", "MATHOP_EMPTY_W_COPY"):
        assert name in ours, name


def test_c_abi_and_ctypes_constants_equal_the_reference():
    c = _defines(os.path.join(ROOT, "include", "clenabled_b200.h"))
    g, m = REF["GRCLBase.h"], REF["clMathOpTypes.h"]
    for suffix in ("COMPLEX", "FLOAT", "INT", "SHORT", "BYTE", "PACKEDXY"):
        assert c["CLB200_DTYPE_" + suffix] == g["DTYPE_" + suffix] == getattr(capi, "DTYPE_" + suffix)
    pairs = {"MULTIPLY": "MULTIPLY", "ADD": "ADD", "SUBTRACT": "SUBTRACT", "COMPLEX_CONJ": "COMPLEX_CONJUGATE",
             "MULTIPLY_CONJ": "MULTIPLY_CONJUGATE", "EMPTY": "EMPTY", "EMPTY_W_COPY": "EMPTY_W_COPY"}
    for ours, theirs in pairs.items():
        assert c["CLB200_OP_" + ours] == m["MATHOP_" + theirs] == getattr(capi, "OP_" + ours)


def test_generated_grc_enums_carry_the_reference_data_type_ids(tmp_path):
    files = {os.path.basename(f): f for f in gen_grc.main(str(tmp_path))}

    def values(name, attr):
        with open(files["b200_%s.block.yml" % name]) as f:
            d = yaml.safe_load(f)
        p = [p for p in d["parameters"] if p["id"] == "type"][0]
        return [int(v) for v in p["option_attributes"][attr]]

    assert values("clXEngine", "data_type") == REF["grc"]["clenabled_clXEngine.type.data_type"]
    assert values("clFFT", "datatype") == REF["grc"]["clenabled_clFFT.type.datatype"]
    assert values("clMultConst", "datatype") == REF["grc"]["clenabled_clMultConst.type.datatype"]


@pytest.mark.skipif(not os.path.isdir("/root/reference/include"), reason="reference tree absent")
def test_fixture_is_what_the_reference_tree_says():
    import importlib.util
    spec = importlib.util.spec_from_file_location("mk", os.path.join(ROOT, "tests", "golden", "make_constants.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    inc = "/root/reference/include/clenabled"
    assert mk.defines(inc + "/GRCLBase.h", ("DTYPE_", "OCLTYPE_", "OCLDEVICESELECTOR_")) == REF["GRCLBase.h"]
    assert mk.defines(inc + "/clMathOpTypes.h", ("MATHOP_",)) == REF["clMathOpTypes.h"]
