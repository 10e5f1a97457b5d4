"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/catgen.h
declares, and refuses to run without a CUDA device (no CPU fallback).  No compute is called here."""
import ctypes
import os
import subprocess
import sys

import pytest

from catgen import lib


def test_library_is_built_and_loads():
    assert os.path.exists(lib.SO_PATH), "build with `make -C cat-generator_b200`"
    L = lib.load()
    assert L.cg_version().decode().startswith("catgen-b200")


def test_every_declared_symbol_is_exported():
    L = lib.load()
    names = lib.declared_symbols()
    assert len(names) > 60, names
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, "declared in include/catgen.h but not exported: %s" % missing


def test_header_is_plain_c():
    # the boundary must be consumable by a C compiler / LuaJIT ffi.cdef: no C++, no torch types
    src = "#include \"catgen.h\"\nint main(void){ cg_step_cfg c; (void)c; return (int)sizeof(cg_model*) == 0; }\n"
    inc = os.path.dirname(lib.HEADER)
    r = subprocess.run(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-I", inc, "-x", "c", "-", "-fsyntax-only"],
                       input=src.encode(), capture_output=True)
    assert r.returncode == 0, r.stderr.decode()
    # signatures only (comments stripped -- they cite reference identifiers such as cutorch.setDevice on purpose)
    import re
    code = re.sub(r"/\*.*?\*/", "", open(lib.HEADER).read(), flags=re.S)
    for banned in ("torch", "at::", "Tensor", "std::", "class ", "template", "&"):
        assert banned not in code, "non-C / torch construct %r in the C-ABI declarations" % banned


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="this check is for the GPU-less build container")
def test_no_gpu_means_loud_failure_not_cpu_fallback():
    code = ("import sys; sys.path.insert(0, %r); from catgen import lib\n"
            "try:\n    lib.init(0)\nexcept lib.CatgenError as e:\n    print('RAISED', e); sys.exit(0)\nsys.exit(3)\n"
            % os.path.dirname(lib.PKG_DIR + "/"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode == 0 and "RAISED" in r.stdout, (r.returncode, r.stdout, r.stderr)
    assert "no CPU fallback" in r.stdout or "no CUDA device" in r.stdout
    # and a compute entry point before init refuses too
    L = lib.load()
    m = ctypes.c_void_p()
    assert L.cg_model_create(ctypes.byref(m), 1, 3, 100, 1) != 0
    assert b"cg_init" in L.cg_last_error()


def _protos(text):
    """{name: (return type, [param types])} for every cg_* prototype in a C declaration text; parameter NAMES are dropped."""
    import re
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"--\[\[.*?\]\]", "", text, flags=re.S)
    out = {}
    for ret, name, args in re.findall(r"([A-Za-z_][A-Za-z0-9_ \*]*?)\s*\b(cg_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", text):
        if ret.strip().startswith("typedef") or ret.strip() in ("return",):
            continue
        types = []
        for a in [x.strip() for x in args.split(",")]:
            if a in ("void", ""):
                continue
            arr = "*" if re.search(r"\[\d*\]\s*$", a) else ""
            a = re.sub(r"\[\d*\]\s*$", "", a).strip()
            m = re.match(r"^(.*?)([A-Za-z_][A-Za-z0-9_]*)$", a)          # strip the trailing identifier (the name)
            t = (m.group(1) if m and m.group(1).strip() else a)
            types.append(re.sub(r"\s+", "", t) + arr)
        out[name] = (re.sub(r"\s+", "", ret), types)
    return out


def test_lua_cdef_matches_the_header():
    """The Lua layer cannot be executed here (no LuaJIT), so at least its ffi.cdef must be the header's ABI verbatim."""
    import re
    lua = open(os.path.join(lib.PKG_DIR, "lua", "catgen_ffi.lua")).read()
    cdef = re.search(r"ffi\.cdef\[\[(.*?)\]\]", lua, flags=re.S).group(1)
    hdr, lu = _protos(open(lib.HEADER).read()), _protos(cdef)
    assert len(lu) >= 25, sorted(lu)
    for name, sig in lu.items():
        assert name in hdr, "%s is bound in Lua but not declared in include/catgen.h" % name
        assert sig == hdr[name], "%s: Lua %s vs header %s" % (name, sig, hdr[name])
    # struct layout of cg_step_cfg: same field order and types
    h_struct = re.sub(r"/\*.*?\*/", "", re.search(r"typedef struct \{(.*?)\} cg_step_cfg;", open(lib.HEADER).read(), flags=re.S).group(1), flags=re.S)
    l_struct = re.search(r"typedef struct \{(.*?)\} cg_step_cfg;", cdef, flags=re.S).group(1)
    def fields(t):   # "int a, b; float c;" -> [("int","a"), ("int","b"), ("float","c")]
        out = []
        for decl in [d.strip() for d in t.split(";") if d.strip()]:
            ty, names = decl.split(None, 1)
            out += [(ty, n.strip()) for n in names.split(",")]
        return out
    assert fields(h_struct) == fields(l_struct) and len(fields(l_struct)) == 13
    # every Lua file says it has not been executed
    for root, _, files in os.walk(os.path.join(lib.PKG_DIR, "lua")):
        for f in files:
            assert "NOT EXECUTED" in open(os.path.join(root, f)).read(), f


def test_lua_surface_covers_what_train_lua_calls():
    """The Lua layer cannot run here; what CAN be checked is that every method the reference's train.lua / adversarial.lua /
    utils/nn_utils.lua invoke on MODEL_G / MODEL_D, every NN_UTILS function train.lua calls, and every module train.lua `require`s
    before building a model exist in cat-generator_b200/lua/ (reference call sites: train.lua:101-107,119-137,147-185,231-261;
    adversarial.lua:84-89,187-197; utils/nn_utils.lua:52,96,334-349,428-462,630)."""
    import re
    lua = os.path.join(lib.PKG_DIR, "lua")
    models = open(os.path.join(lua, "models.lua")).read()
    for m in ("forward", "backward", "getParameters", "training", "evaluate", "zeroGradParameters", "clone", "cuda", "float", "listModules",
              "clearState", "__tostring", "write", "read"):
        assert re.search(r"function Net:%s\b" % re.escape(m), models), "catgen.Net lacks :%s()" % m
    assert "self.modules = {self}" in models and "self.gradInput" in models       # MODEL_D.modules[1].gradInput (adversarial.lua:193)
    utils = open(os.path.join(lua, "utils", "nn_utils.lua")).read()
    for f in ("createNoiseInputs", "createImagesFromNoise", "createImages", "sortImagesByPrediction", "switchToTrainingMode", "switchToEvaluationMode",
              "prepareNetworkForSave", "getNumberOfParameters", "activateCuda", "visualizeProgress", "rateWithV"):
        assert re.search(r"function nn_utils\.%s\b" % f, utils), "utils/nn_utils.lua lacks %s" % f
    for req, path in (("cutorch", "rocks/cutorch.lua"), ("cunn", "rocks/cunn.lua"), ("dpnn", "rocks/dpnn.lua"), ("stn", "rocks/stn.lua"), ("cudnn", "rocks/cudnn.lua"),
                      ("LeakyReLU", "LeakyReLU.lua"), ("layers.cudnnSpatialConvolutionUpsample", "layers/cudnnSpatialConvolutionUpsample.lua"),
                      ("adversarial", "adversarial.lua"), ("models", "models.lua"), ("utils.nn_utils", "utils/nn_utils.lua")):
        assert os.path.exists(os.path.join(lua, path)), "train.lua requires %r: %s is missing" % (req, path)
    cu = open(os.path.join(lua, "layers", "cudnnSpatialConvolutionUpsample.lua")).read()
    assert "torch.class('cudnn.SpatialConvolutionUpsample'" in cu and "accUpdateGradParameters" in cu
    adv = open(os.path.join(lua, "adversarial.lua")).read()
    assert "thisB - thisB % 2" in adv and "maxAccuracyD <= 1" in adv and "syncToHost" in adv
    # every C function the Lua files call is declared in the cdef (and therefore, by the test above, in the header)
    cdef = re.search(r"ffi\.cdef\[\[(.*?)\]\]", open(os.path.join(lua, "catgen_ffi.lua")).read(), flags=re.S).group(1)
    declared = set(re.findall(r"\b(cg_[a-z0-9_]+)\s*\(", cdef))
    for root, _, files in os.walk(lua):
        for f in files:
            for name in re.findall(r"cg\.lib\.(cg_[a-z0-9_]+)", open(os.path.join(root, f)).read()):
                assert name in declared, "%s calls %s, which catgen_ffi.lua does not declare" % (f, name)
