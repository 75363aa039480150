"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol the header declares,
struct layouts agree between C and ctypes, and the product fails loudly without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "motion_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(md_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(capi):
    lib = capi.lib()
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), "libmotion_b200.so does not export %s" % s
    assert sorted(capi.SYMBOLS) == syms


def test_struct_layouts_match_c(capi, tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "motion_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(md_config),sizeof(md_frames),sizeof(md_outputs),sizeof(md_stats),offsetof(md_config,lk_eps),'
                   'offsetof(md_config,ransac_thresh),offsetof(md_config,vf_rho),sizeof(md_live_params),sizeof(md_live_result),'
                   'offsetof(md_live_params,distance_threshold),offsetof(md_live_result,traj),offsetof(md_stats,lk_iterations));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(capi.MdConfig), C.sizeof(capi.MdFrames), C.sizeof(capi.MdOutputs), C.sizeof(capi.MdStats),
            capi.MdConfig.lk_eps.offset, capi.MdConfig.ransac_thresh.offset, capi.MdConfig.vf_rho.offset,
            C.sizeof(capi.MdLiveParams), C.sizeof(capi.MdLiveResult), capi.MdLiveParams.distance_threshold.offset,
            capi.MdLiveResult.traj.offset, capi.MdStats.lk_iterations.offset]
    assert got == want


def test_defaults_are_the_reference_constants(capi):
    cfg = capi.default_config()
    # optical_flow_calculator.cpp:40-44,71,127 ; outlier_detector.cpp:250 ; cpp:422-429
    assert (cfg.lk_win, cfg.lk_max_level, cfg.lk_max_iters) == (40, 5, 10)
    assert abs(cfg.lk_eps - 0.03) < 1e-12 and abs(cfg.lk_min_eig - 0.001) < 1e-9
    assert cfg.diff_threshold == 190 and cfg.morph == 1 and cfg.ransac_iters == 50
    assert (cfg.vf_max_level, cfg.vf_start_level, cfg.vf_n1, cfg.vf_n2) == (4, 0, 2, 2)
    assert abs(cfg.vf_rho - 2.8) < 1e-6 and cfg.vf_alpha == 1400.0 and cfg.vf_sigma == 1.5
    assert cfg.pixel_step == 10 and cfg.min_vector_size == 1.0


def test_no_cpu_fallback_without_gpu(capi):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.MotionB200Error) as e:
        capi.Context(width=64, height=64)
    assert e.value.code == -2


def test_invalid_arguments_return_status_codes(capi):
    lib = capi.lib()
    assert lib.md_config_default(None) == -1
    cfg = capi.default_config(width=4, height=4)
    h = C.c_void_p()
    assert lib.md_create(C.byref(cfg), 0, C.byref(h)) == -1
    assert lib.md_destroy(None) == -1
    assert lib.md_grid_size(None) == -1


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "motion_detection_b200")
    # the product may MENTION the oracle in comments (operation-order notes); it must never import, include or link it
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            path = os.path.join(dp, f)
            if f.endswith(".py"):
                for line in open(path):
                    s = line.strip()
                    assert not re.match(r"(from\s+\.*oracle|import\s+oracle|from\s+\S*\boracle\b\S*\s+import)", s), (f, s)
                    assert "libmd_oracle" not in s, (f, s)
            elif f.endswith((".cu", ".cuh", ".h", ".cpp")) or f == "Makefile":
                for line in open(path):
                    s = line.strip()
                    if s.startswith("#include"):
                        assert "oracle" not in s, (f, s)
                    assert "md_oracle.h" not in s and "lmd_oracle" not in s and "libmd_oracle" not in s, (f, s)
    # and the shared library has no dependency on it
    out = subprocess.run(["ldd", os.path.join(pkg, "lib", "libmotion_b200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_unpack_mask_host_matches_numpy():
    """md_unpack_mask_host is pure host code (runs without a GPU): packed mask bits, LSB first, -> bytes 0 / 255."""
    import numpy as np
    from motion_detection_b200 import capi
    rng = np.random.default_rng(3)
    for w in (1, 7, 8, 9, 64, 333, 1920):
        pitch = (w + 7) // 8 + 2
        bits = rng.integers(0, 256, (2, 5, pitch), dtype=np.uint8)
        ref = np.unpackbits(bits, axis=-1, bitorder="little")[..., :w] * np.uint8(255)
        assert np.array_equal(capi.unpack_mask(bits, w), ref), w
    one = np.zeros((1, 1), np.uint8)
    out = np.zeros((1, 16), np.uint8)
    assert capi.lib().md_unpack_mask_host(capi._ptr(one), 1, 16, 1, capi._ptr(out), 16) == -1      # bits_pitch too small: MD_ERR_INVALID
    assert capi.lib().md_unpack_mask_host(None, 1, 8, 1, capi._ptr(out), 16) == -1


def test_pipeline_snippet_of_integration_md_compiles_as_c99(tmp_path):
    """The two-contexts loop INTEGRATION.md shows (MD_MEM_HOST_ASYNC + md_set_pair_index) is valid C against the header."""
    import re
    import subprocess
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    m = re.search(r"```c\n(md_ctx \*lane\[2\];.*?)```", text, re.S)
    assert m, "pipeline snippet not found in INTEGRATION.md"
    body = m.group(1).replace("md_ctx *lane[2];", "", 1)
    src = tmp_path / "snip.c"
    src.write_text(
        "#include <stddef.h>\n#include \"motion_b200.h\"\n"
        "static void consume(md_outputs o) { (void)o; }\n"
        "int run(md_ctx *lane[2], const uint8_t *clip, int B, int pitch, long long frame_bytes, md_outputs out[2], int nbatches)\n{\n"
        + body + "    return 0;\n}\n")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
