"""CPU tests of the host side: the C-ABI library loads and exports every symbol include/maskbit_b200.h declares (no
compute calls without a GPU), configs / schedule tables / checkpoint plumbing mirror the reference, and the product
path fails loudly without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

import maskbit_b200
from maskbit_b200 import ConvVQModel, LFQBert, _lib, load_config, sample, sampler_kwargs
from maskbit_b200.masking import get_masking_ratio, step_tables
from maskbit_b200.weights import conv_vq_spec, lfq_bert_spec
from oracle import maskbit_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "maskbit_b200.h")).read()
    declared = set(re.findall(r"^\s*(?:const\s+char\*|int64_t|int|void)\s+(mb_\w+)\s*\(", header, flags=re.M))
    assert declared, "no declarations parsed from the header"
    assert declared == set(_lib.SYMBOLS), f"header / ctypes table mismatch: {declared ^ set(_lib.SYMBOLS)}"
    L = _lib.lib()
    for name in declared:
        assert getattr(L, name) is not None
    assert L.mb_version().decode().endswith("sm_100a")


def test_abi_argument_errors_without_gpu():
    L = _lib.lib()
    assert L.mb_create(None, None) == -1
    assert b"null argument" in L.mb_last_error()
    cfg = _lib.MBConfig()
    cfg.hidden_dim, cfg.heads, cfg.depth, cfg.mlp_dim, cfg.token_bits, cfg.codebook_splits, cfg.seq_len = 768, 12, 2, 3072, 12, 2, 256
    h = ctypes.c_void_p()
    assert L.mb_create(ctypes.byref(cfg), ctypes.byref(h)) == -1       # hidden_dim 768 unsupported
    assert b"hidden_dim" in L.mb_last_error()
    cfg.hidden_dim, cfg.heads = 1024, 8                                 # head dim 128: the attention kernels are built for 64
    assert L.mb_create(ctypes.byref(cfg), ctypes.byref(h)) == -1
    assert b"head dim" in L.mb_last_error()
    assert L.mb_launch_count(None) == 0


def test_struct_layouts_match_header():
    """ctypes mirrors of mb_config / mb_select_args / mb_sample_args have the C layout (sizes from a gcc probe)."""
    import subprocess
    import tempfile
    src = '#include <stdio.h>\n#include "maskbit_b200.h"\nint main(){printf("%zu %zu %zu\\n", sizeof(mb_config), sizeof(mb_select_args), sizeof(mb_sample_args));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "p.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "p"), os.path.join(d, "p.c")])
        sizes = [int(x) for x in subprocess.check_output([os.path.join(d, "p")]).split()]
    assert sizes == [ctypes.sizeof(_lib.MBConfig), ctypes.sizeof(_lib.MBSelectArgs), ctypes.sizeof(_lib.MBSampleArgs)]


def test_product_path_has_no_cpu_fallback():
    cfg = load_config("maskbit_generator_12bit")
    kw = sampler_kwargs(cfg)
    tok = ConvVQModel(cfg.model.vq_model, legacy=False)
    gen = LFQBert(img_size=256, hidden_dim=1024, codebook_size=4096, codebook_splits=2, depth=24, heads=16, mlp_dim=4096,
                  dropout=0.0, use_prenorm=False, input_stride=16)
    assert gen.device.type == "cpu"
    with pytest.raises(_lib.MaskbitError, match="no CPU fallback"):
        gen(torch.zeros((1, 256, 2), dtype=torch.int64), torch.zeros(1, dtype=torch.int64))
    with pytest.raises(_lib.MaskbitError, match="no CPU fallback"):
        tok.decode_tokens(torch.zeros((1, 256), dtype=torch.int64))
    with pytest.raises(_lib.MaskbitError, match="no CPU fallback"):
        sample(gen, tok, num_samples=1, labels=torch.zeros(1, dtype=torch.long), **kw)
    with pytest.raises(NotImplementedError):
        gen.train()
    # nothing under the package imports the oracle
    for root, _, files in os.walk(os.path.join(ROOT, "maskbit_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".sh")):
                for line in open(os.path.join(root, f)):
                    if re.match(r"\s*(from|import|#include)\b", line) or "CDLL" in line or "dlopen" in line:
                        assert "oracle" not in line, f"{f}: {line}"


def test_model_attributes_mirror_reference():
    # bert.py:360-377
    for bits, v in [(10, 32), (12, 64), (14, 128), (16, 256), (18, 512)]:
        gen = LFQBert(img_size=256, hidden_dim=1024, codebook_size=2 ** bits, codebook_splits=2, depth=24, heads=16, mlp_dim=4096,
                      dropout=0.1, input_stride=16)
        assert gen.seq_len == 256 and gen.splits == 2 and gen.get_group_splits() == 2
        assert gen.effective_codebook_size == v and gen.mask_token == v and gen.bits == bits
        assert gen.drop_label == 1000 and gen.dtype == torch.float32
    with pytest.raises(NotImplementedError):
        ConvVQModel(load_config("maskbit_generator_12bit").model.vq_model, legacy=True)


@pytest.mark.parametrize("name", ["maskbit_generator_10bit", "maskbit_generator_12bit", "maskbit_generator_14bit",
                                  "maskbit_generator_14bit_128steps", "maskbit_generator_14bit_256steps",
                                  "maskbit_generator_16bit", "maskbit_generator_18bit"])
def test_configs_and_kwarg_mapping(name):
    cfg = load_config(name)
    kw = sampler_kwargs(cfg)                                   # eval_maskbit.py:74-80,114-132
    bits = cfg.model.vq_model.token_size
    assert kw["codebook_size"] == 2 ** bits and kw["mask_token"] == 2 ** (bits // 2) and kw["patch_size"] == 16
    assert kw["codebook_splits"] == 2 and kw["guidance_annealing"] == "cosine" and kw["mask_schedule_strategy"] == "arccos"
    assert cfg.model.vq_model.get("num_res_blocks_decoder", 7) == 7     # OmegaConf-style .get (autoencoder.py:371)
    with pytest.raises(ValueError):
        load_config("does_not_exist")


def test_schedule_tables_equal_oracle():
    for mode in ["root", "square", "cosine", "arccos", "linear"]:
        for t in (8, 64):
            for i in range(t):
                assert torch.equal(get_masking_ratio((i + 1) / t, mode), O.get_masking_ratio((i + 1) / t, mode))
    for ann in ["none", "linear", "cosine"]:
        scale, temp, omp, mask_len = step_tables(64, 512, softmax_temperature=1.0, mask_schedule_strategy="arccos",
                                                 guidance_scale=7.1, guidance_annealing=ann, scale_pow=3.0,
                                                 use_sampling_annealing=False)
        for i in range(64):
            want = O.guidance_scale_at(i, 64, 7.1, ann, 3.0)
            assert scale[i] == float(torch.as_tensor(want, dtype=torch.float32).reshape(-1)[0])
        assert temp == [1.0] * 64 and mask_len[-1] == 0.0 and mask_len[0] == 506.0
    _, temp, _, _ = step_tables(4, 512, softmax_temperature=1.0, mask_schedule_strategy="linear", guidance_scale=0.0,
                                guidance_annealing="none", scale_pow=1.0, use_sampling_annealing=True)
    assert temp == [0.5 + 0.8 * (1 - (i + 1) / 4) for i in range(4)]          # sampling.py:103-104


def test_state_dict_layout_and_save_load_roundtrip(tmp_path):
    """Key names / shapes are the reference's and save_pretrained -> load_pretrained round-trips."""
    spec = {n: s for n, s, _ in lfq_bert_spec()}
    assert spec["pos_emb"] == (1, 257, 1024) and spec["transformer.layers.23.0.mha.in_proj_weight"] == (3072, 1024)
    assert spec["prediction_layer.weight"] == (128, 1024) and spec["input_proj.weight"] == (1024, 12)
    assert sum(int(torch.tensor(s).prod()) for n, s in spec.items() if n != "bits_to_indices") == 304_795_776   # 304.80 M parameters
    dspec = {n: s for n, s, _ in conv_vq_spec()}
    assert dspec["decoder.conv_in.weight"] == (512, 12, 3, 3) and dspec["decoder.up.1.res_blocks.0.nin_shortcut.weight"] == (256, 256, 1, 1)
    assert "decoder.up.4.upsample_conv.weight" not in dspec and dspec["quantize.codebook"] == (4096, 12)
    cfg = load_config("maskbit_generator_12bit")
    tok = ConvVQModel(cfg.model.vq_model, legacy=False)
    tok.save_pretrained(str(tmp_path / "tok"))
    tok2 = ConvVQModel(cfg.model.vq_model, legacy=False)
    tok2.load_pretrained(str(tmp_path / "tok"))
    a, b = tok.state_dict(), tok2.state_dict()
    assert list(a) == list(b) and all(torch.equal(a[k], b[k]) for k in a)
    bad = dict(a)
    bad.pop("decoder.conv_out.bias")
    with pytest.raises(RuntimeError, match="Missing key"):
        tok2.load_state_dict(bad)
    with pytest.raises(maskbit_b200._lib.MaskbitError):        # integer helpers are device kernels too: no CPU fallback
        maskbit_b200.split_factorized_tokens(torch.tensor([[37]]), 4096, 2)


def test_eval_driver_label_schedule():
    """eval_maskbit.py:107-112: randperm(1000) repeated, cut into consecutive batches -- every block of 1000 holds each class once."""
    from maskbit_b200.eval_driver import label_schedule
    lab = label_schedule(2500, label_seed=5)
    assert lab.dtype == torch.int64 and lab.shape == (2500,)
    assert sorted(lab[:1000].tolist()) == list(range(1000)) and torch.equal(lab[:1000], lab[1000:2000]) and torch.equal(lab[:500], lab[2000:])
    assert torch.equal(lab, label_schedule(2500, label_seed=5)) and not torch.equal(lab, label_schedule(2500, label_seed=6))


def test_bench_flop_accounting_matches_survey():
    """bench.py counts the algorithmic FLOPs SURVEY.md 8(d) states (2*MAC, CFG on every step, all 257 rows through the head)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert bench.f_fwd(12) == 162_328_313_856 and bench.f_fwd(14) == 162_396_733_440
    assert abs(bench.f_img(12, 64) - 20.964e12) < 0.001e12 and abs(bench.f_img(12, 8) - 2.783e12) < 0.001e12
    assert abs(2 * 256 * bench.f_fwd(12) - 83.11e12) < 0.01e12          # transformer-step roofline numerator at B = 256


def test_bench_other_kernel_rooflines_accounting():
    """bench.py's per-kernel rooflines: with every GEMM class given the time its FLOPs take at exactly 1000 TFLOP/s and attention the
    time its algorithmic bytes take at 1000 GB/s, the helper must report those rates; per sequence and layer the four GEMMs add up
    to the 6.47 GFLOP DESIGN.md section 4 states."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    S, D, MLP, DEPTH = bench.S, bench.D, bench.MLP, bench.DEPTH
    per_seq = 2 * S * (3 * D * D + D * D + 2 * D * MLP)
    assert abs(per_seq - 6.47e9) < 0.01e9
    seqs, steps = [512] * 61 + [256] * 3, 2
    rows = steps * sum(n * S for n in seqs)
    prof = {"gemm_qkv": (2.0 * rows * 3 * D * D * DEPTH / 1e12, 10), "gemm_out": (2.0 * rows * D * D * DEPTH / 1e12, 10),
            "gemm_down": (2.0 * rows * D * MLP * DEPTH / 1e12, 10), "attention": (rows * 4 * D * 2.0 * DEPTH / 1e9, 10)}
    peaks = dict(bf16_sustained=1358.3, bf16_burst=1601.2, hbm=6549.8, source="measured")
    out = bench.other_kernel_rooflines(prof, steps, seqs, peaks)
    for k in ("gemm_qkv", "gemm_out", "gemm_down"):
        assert abs(out[k]["achieved"] - 1000.0) < 1e-6 and out[k]["unit"] == "TFLOP/s" and abs(out[k]["frac"] - 1000.0 / 1358.3) < 1e-9
    assert abs(out["attention"]["achieved"] - 1000.0) < 1e-6 and out["attention"]["bound"] == "hbm"
    assert "error" not in bench.safe_other_kernel_rooflines(prof, steps, seqs, peaks)
