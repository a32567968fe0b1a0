"""CPU: the C-ABI library loads, exports every symbol include/swarm_b200.h declares, the ctypes
structs have the C layout (checked with gcc), and argument validation works without a GPU."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "swarm_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(swarm_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound(pkg):
    nat = pkg._native
    lib = nat.load()
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), "libswarm_b200.so does not export %s" % n
        assert n in nat.SYMBOLS, "ctypes binding misses %s" % n
    assert sorted(nat.SYMBOLS) == names
    assert lib.swarm_abi_version() == 3
    assert lib.swarm_strerror(0) == b"ok" and b"NULL" in lib.swarm_strerror(-1)


def test_ctypes_structs_match_c_layout(pkg, tmp_path):
    nat = pkg._native
    structs = {"SwarmParams": nat.SwarmParams, "SwarmState": nat.SwarmState,
               "SwarmInjectedDraws": nat.SwarmInjectedDraws, "SwarmStepIO": nat.SwarmStepIO}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "swarm_b200.h"', "int main(void){"]
    for sname, cls in structs.items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (sname, sname))
        for f, _ in cls._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (sname, f, sname, f))
    lines.append("return 0;}")
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    for sname, cls in structs.items():
        assert int(out[sname]) == ctypes.sizeof(cls), sname
        for f, _ in cls._fields_:
            assert int(out["%s.%s" % (sname, f)]) == getattr(cls, f).offset, (sname, f)


def _params(nat, **kw):
    base = dict(n_envs=4, n_locusts=80, n_agents=10, grid_size=84, n_burn_in=10, max_episode_steps=128,
                math_mode=0, tuning=0, noise=1e-4, gravity=-1.0, wind=1.0, F=0.5, L=10.0, dt=0.05,
                box_width=3.0, box_height=3.0, seed=1, env_id_offset=0)
    base.update(kw)
    return nat.SwarmParams(**base)


def test_validation_without_gpu(pkg):
    nat = pkg._native
    lib = nat.load()
    assert lib.swarm_validate(ctypes.byref(_params(nat))) == 0
    assert lib.swarm_validate(None) == -1
    for bad in (dict(n_envs=0), dict(n_locusts=0), dict(n_locusts=4096), dict(n_agents=0), dict(grid_size=256),
                dict(grid_size=1), dict(env_id_offset=-1), dict(env_id_offset=2 ** 32)):
        assert lib.swarm_validate(ctypes.byref(_params(nat, **bad))) == -2, bad
    assert lib.swarm_validate(ctypes.byref(_params(nat, math_mode=7))) == -4
    # NULL buffers are rejected before any CUDA call
    p = _params(nat)
    assert lib.swarm_reset(ctypes.byref(p), None, None, None, None) == -1
    st = nat.SwarmState()
    assert lib.swarm_step(ctypes.byref(p), ctypes.byref(st), None, None, None) == -1
    assert lib.swarm_rasterize(ctypes.byref(p), None, None, None, None, None, None) == -1
    assert lib.swarm_clip_actions(None, 4, 1.0, None) == -1
    with pytest.raises(nat.SwarmNativeError):
        nat.check(-2, "x")


def test_product_package_does_not_import_oracle():
    """The oracle is test infrastructure; the product must not route through it."""
    pkgdir = os.path.join(ROOT, "golds-rl-gym_b200")
    for dp, _, files in os.walk(pkgdir):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
    code = "import sys; sys.path.insert(0, %r); import golds_rl_gym_b200 as p; p.submodule('envs.multiagent'); " \
           "p.submodule('agents.paac.runners'); assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules)" % ROOT
    subprocess.check_call([sys.executable, "-c", code])


def test_no_cpu_fallback_without_cuda(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    m = pkg.submodule("envs.multiagent")
    with pytest.raises(pkg.SwarmNativeError):
        m.BatchedSwarmEnv(4)


def test_torch_extension_loads_and_registers_ops(pkg):
    """The C ABI as a PyTorch extension (csrc/torch_binding.cpp): builds on CPU, registers
    torch.ops.swarm_b200.*, validates arguments before touching the GPU."""
    import torch
    ops = pkg.load_torch_ops()
    assert int(ops.abi_version()) == 3
    for name in ("step", "reset", "rasterize", "expand_obs", "clip_actions", "forces"):
        assert hasattr(ops, name)
    with pytest.raises(RuntimeError):
        ops.clip_actions(torch.zeros(4, 2), 1.0)                       # CPU tensor: refused, no CPU path
    nat = pkg._native
    p = nat.SwarmParams(n_envs=2, n_locusts=8, n_agents=10, grid_size=84, n_burn_in=10, max_episode_steps=128,
                        math_mode=0, tuning=0, noise=1e-4, gravity=-1.0, wind=1.0, F=0.5, L=10.0, dt=0.05,
                        box_width=3.0, box_height=3.0, seed=0, env_id_offset=0)
    blob = nat.params_blob(p)
    assert blob.dtype == torch.uint8 and blob.numel() == 112
    with pytest.raises(RuntimeError):
        ops.forces(blob[:10].clone(), torch.zeros(2, 8, 2, dtype=torch.float64), torch.zeros(2, 10, 2, dtype=torch.float64),
                   None, None)                                         # wrong params size
