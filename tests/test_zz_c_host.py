"""The C ABI from a compiled host: examples/c_host/hfg_host.c is plain C11 over include/hfg.h
(no Python, no torch, no CUDA headers) and links libhfg_b200.so directly -- the binding a
compiled-language maintainer would write for HiFiGANGenerator (reference models/hifigan.py:149-261:
constructor, load_state_dict, remove_weight_norm, forward).

CPU: it compiles against the header with -Wall -Wextra -Werror, answers the host-only entry points and
FAILS LOUDLY in hfg_create (HFG_ERR_CUDA) -- there is no CPU path to fall back to.
GPU: fed the raw tensors of a golden case of the live reference, its waveform equals the Python mirror's
(same library entry point: bit for bit) and the golden within the mode's tolerance.
"""
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

from conftest import ROOT, case_inputs, load_golden

LIBDIR = os.path.join(ROOT, "tts-sambert_hifigan_b200", "lib")
SRC = os.path.join(ROOT, "examples", "c_host", "hfg_host.c")
MODE_ID = {"fp32": 0, "tf32": 1, "bf16": 2, "fp16": 3}
TOL = {"fp32": 5e-7, "tf32": 1.8e-4, "fp16": 2.5e-4, "bf16": 1.8e-3}      # tests/test_parity_gpu.py


@pytest.fixture(scope="module")
def host_binary(tmp_path_factory):
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    if not os.path.exists(os.path.join(LIBDIR, "libhfg_b200.so")):
        pytest.fail("libhfg_b200.so is not built: run __graft_entry__.build()")
    exe = str(tmp_path_factory.mktemp("c_host") / "hfg_host")
    cmd = ["gcc", "-std=c11", "-O2", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), SRC, "-o", exe,
           "-L", LIBDIR, "-lhfg_b200", "-Wl,-rpath," + LIBDIR]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    return exe


def cfg_line(cfg):
    rates, kernels = cfg["upsample_rates"], cfg["upsample_kernel_sizes"]
    pairs = list(zip(cfg["resblock_kernel_sizes"], cfg["resblock_dilation_sizes"]))
    toks = ["cfg", cfg["n_mels"], cfg["upsample_initial_channel"], len(rates), *rates, *kernels, len(pairs)]
    for k, dils in pairs:
        toks += [k, len(dils), *dils]
    return " ".join(str(t) for t in toks)


def export_case(dirname, cfg, sd, mel):
    """What any exporter of a reference checkpoint would write: raw fp32 tensors in the reference's layout."""
    lines = [cfg_line(cfg)]
    np.ascontiguousarray(mel, np.float32).tofile(os.path.join(dirname, "mel.bin"))
    lines.append(f"mel {mel.shape[0]} {mel.shape[2]} mel.bin")
    for i, (key, w) in enumerate(sd.items()):
        w = np.ascontiguousarray(w, np.float32)
        w.tofile(os.path.join(dirname, f"w{i}.bin"))
        lines.append(f"w {key} {w.ndim} {' '.join(str(d) for d in w.shape)} w{i}.bin")
    with open(os.path.join(dirname, "manifest.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")


@pytest.mark.parametrize("header", ["hfg.h", "hfg_ard.h", "hfg_mel.h"])
def test_headers_are_plain_c_and_cxx(header, tmp_path):
    """No torch / C++ types in any signature: every header is a translation unit of its own in C99 and in C++11."""
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    for lang, std, cc in (("c", "-std=c99", "gcc"), ("c++", "-std=c++11", "g++")):
        out = subprocess.run([cc, std, "-pedantic", "-Wall", "-Werror", "-x", lang, "-fsyntax-only", "-I", os.path.join(ROOT, "include"), "-"],
                             input=f'#include "{header}"\n', capture_output=True, text=True)
        assert out.returncode == 0, (lang, out.stderr)


@pytest.mark.skipif(torch.cuda.is_available(), reason="the no-device behaviour needs a box without a GPU")
def test_c_host_fails_loudly_without_a_gpu(host_binary, manifest, tmp_path):
    cfg, _, _ = case_inputs(manifest, "default_b2_t24")
    with open(tmp_path / "manifest.txt", "w") as f:
        f.write(cfg_line(cfg) + "\n")
    out = subprocess.run([host_binary, str(tmp_path)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 3, (out.stdout, out.stderr)
    assert "abi 3 (header 3) receptive_radius 13 rc 0" in out.stdout          # host-only entry points still answer
    assert "hfg_create failed: -2 HFG_ERR_CUDA" in out.stderr and "no CPU fallback" in out.stderr
    assert not os.path.exists(tmp_path / "wav.bin")


@pytest.mark.gpu
@pytest.mark.parametrize("name,mode", [("default_b2_t24", "tf32"), ("default_weightnorm_b1_t16", "fp32"),
                                       ("small_custom_b3_t33", "bf16")])
def test_c_host_equals_python_mirror_and_golden(host_binary, manifest, name, mode, tmp_path):
    import tts_sambert_hifigan_b200 as pkg
    cfg, sd, mel = case_inputs(manifest, name)
    export_case(str(tmp_path), cfg, sd, mel)
    out = subprocess.run([host_binary, str(tmp_path), str(MODE_ID[mode])], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, (out.stdout, out.stderr)
    ok = [l for l in out.stdout.splitlines() if l.startswith("ok ")][0].split()
    g = load_golden(name)["wav"]
    assert int(ok[1]) == g.shape[0] and int(ok[2]) == g.shape[2] and int(ok[3]) > 0 and int(ok[4]) == len(sd)
    wav = np.fromfile(tmp_path / "wav.bin", np.float32).reshape(g.shape)
    err = float(np.abs(wav - g).max())
    print(f"c_host {name}[{mode}]: max-abs vs golden {err:.3e}, launches {ok[3]}")
    assert err <= TOL[mode]
    # the Python mirror on CPU tensors goes through the same entry point (hfg_forward_host*): identical bits
    gen = pkg.HiFiGANGenerator(**cfg, mode=mode).to("cuda:0")
    gen.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    with torch.no_grad():
        ref = gen(torch.from_numpy(mel)).numpy()
    assert np.array_equal(ref, wav)
