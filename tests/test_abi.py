"""CPU-side checks of the drop-in boundary: the library loads, exports every symbol include/cadm_b200.h declares,
the ctypes struct mirrors the C struct, and the product fails loudly without a GPU (no CPU fallback)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "cadm_b200.h")


@pytest.fixture(scope="module")
def lib():
    from cadm_b200 import build
    build.build()
    from cadm_b200 import _lib
    return _lib.load()


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cadm_[a-z_0-9]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    from cadm_b200 import _lib
    syms = _declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in cadm_b200.h but not exported"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature"
    assert set(_lib.SIGNATURES) == set(syms)
    assert lib.cadm_abi_version() == 1


def test_config_struct_matches_header(tmp_path):
    """sizeof / field offsets of CadmConfig as seen by gcc equal the ctypes mirror."""
    from cadm_b200._lib import CadmConfig
    fields = [f[0] for f in CadmConfig._fields_]
    prog = '#include <stdio.h>\n#include <stddef.h>\n#include "cadm_b200.h"\nint main(){printf("%zu", sizeof(CadmConfig));' + \
        "".join(f'printf(" %zu", offsetof(CadmConfig, {f}));' for f in fields) + "return 0;}"
    c = tmp_path / "t.c"
    c.write_text(prog)
    exe = tmp_path / "t"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(c), "-o", str(exe)])
    vals = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    assert vals[0] == ctypes.sizeof(CadmConfig)
    assert vals[1:] == [getattr(CadmConfig, f).offset for f in fields]


def test_train_config_struct_matches_header(tmp_path):
    """sizeof / field offsets of CadmTrainConfig (fit() on the device) as seen by gcc equal the ctypes mirror."""
    from cadm_b200._lib import CadmTrainConfig
    fields = [f[0] for f in CadmTrainConfig._fields_]
    prog = '#include <stdio.h>\n#include <stddef.h>\n#include "cadm_b200.h"\nint main(){printf("%zu", sizeof(CadmTrainConfig));' + \
        "".join(f'printf(" %zu", offsetof(CadmTrainConfig, {f}));' for f in fields) + "return 0;}"
    c = tmp_path / "t.c"
    c.write_text(prog)
    exe = tmp_path / "t"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(c), "-o", str(exe)])
    vals = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    assert vals[0] == ctypes.sizeof(CadmTrainConfig)
    assert vals[1:] == [getattr(CadmTrainConfig, f).offset for f in fields]


def test_trainer_has_no_cpu_fallback(lib):
    """cadm_train_create validates its configuration and, without a CUDA device, fails instead of training on the host."""
    import torch
    from cadm_b200._lib import CadmTrainConfig
    cfg = CadmTrainConfig()
    h = ctypes.c_void_p()
    assert lib.cadm_train_create(ctypes.byref(cfg), ctypes.byref(h)) == -1         # struct_size not set
    assert b"struct_size" in lib.cadm_train_last_error(None)
    cfg.struct_size = ctypes.sizeof(CadmTrainConfig)
    cfg.obs_dim, cfg.proc_obs_dim, cfg.act_dim, cfg.hidden, cfg.n_hidden, cfg.ensemble = 18, 18, 6, 200, 4, 5
    cfg.has_back = 1                                                               # backward model without a context encoder
    assert lib.cadm_train_create(ctypes.byref(cfg), ctypes.byref(h)) == -1
    cfg.has_back = 0
    if not torch.cuda.is_available():
        assert lib.cadm_train_create(ctypes.byref(cfg), ctypes.byref(h)) == -2     # CADM_ERR_CUDA
        assert b"no CPU fallback" in lib.cadm_train_last_error(None)


def test_create_rejects_bad_config(lib):
    from cadm_b200.engine import PlannerConfig
    cfg = PlannerConfig(particles=7, ensemble=5).to_c()
    h = ctypes.c_void_p()
    assert lib.cadm_plan_create(ctypes.byref(cfg), ctypes.byref(h)) != 0
    assert b"particles" in lib.cadm_last_error(None)
    cfg = PlannerConfig().to_c()
    cfg.struct_size = 4
    assert lib.cadm_plan_create(ctypes.byref(cfg), ctypes.byref(h)) != 0
    assert b"ABI" in lib.cadm_last_error(None)


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from cadm_b200._lib import CadmError
    from cadm_b200.synth import build_model
    with pytest.raises(CadmError):
        build_model("C2")
    from cadm_b200.engine import PlannerConfig
    cfg = PlannerConfig().to_c()
    h = ctypes.c_void_p()
    assert lib.cadm_plan_create(ctypes.byref(cfg), ctypes.byref(h)) == -2          # CADM_ERR_CUDA
    assert b"no CPU fallback" in lib.cadm_last_error(None)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under cadm_b200/ may import it."""
    pkg = os.path.join(ROOT, "cadm_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), os.path.join(dirpath, f)
    code = "import sys; import cadm_b200, cadm_b200.engine, cadm_b200.parallel, cadm_b200.synth; " \
           "import cadm_b200.dynamics.mlp_cadm_ensemble_cem_dynamics, cadm_b200.policies.mpc_controller; " \
           "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules), 'oracle imported'"
    subprocess.check_call([sys.executable, "-c", code], cwd=ROOT)
