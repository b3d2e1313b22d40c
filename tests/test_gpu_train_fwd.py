"""GPU parity of the training step's forward half (reference scripts/train_maskbit.py:362-380) through the
reference-facing mirrors (maskbit_b200.split_factorized_tokens / get_mask_tokens / MLMLoss -> C ABI) against the reference's own
outputs (tests/golden/train_fwd.npz) and the oracle."""
import os

import numpy as np
import pytest
import torch

import maskbit_b200
from maskbit_b200 import MLMLoss, get_mask_tokens, split_factorized_tokens
from oracle import maskbit_oracle as O

pytestmark = pytest.mark.gpu


def _inputs(v, b, seed):
    g = torch.Generator().manual_seed(seed)
    full = torch.randint(0, v * v, (b, 256), generator=g)
    logits = torch.randn((b, 256, 2, v), generator=g) * 2.0
    return full, logits


def test_training_forward_half_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "train_fwd.npz"))
    for v, b in ((64, 6), (128, 3)):
        full, logits = _inputs(v, b, seed=900 + v)
        tok = split_factorized_tokens(full.cuda(), v * v, 2)
        assert tok.dtype == torch.int64 and np.array_equal(tok.cpu().numpy(), g[f"v{v}_split"].astype(np.int64))
        assert torch.equal(maskbit_b200.combine_factorized_tokens(tok, v * v, 2).long().cpu(), full)     # round trip
        logits = logits.clone()
        logits.scatter_add_(-1, tok.cpu().unsqueeze(-1), torch.full(tuple(tok.shape) + (1,), 2.5))
        for mode in ("arccos", "linear", "square", "cosine"):
            torch.manual_seed(77)
            masked, mask = get_mask_tokens(tok, v, mode=mode, min_masking_ratio=0.1 if mode == "square" else 0.0)
            assert masked.is_cuda and mask.dtype == torch.bool
            assert np.array_equal(masked.cpu().numpy(), g[f"v{v}_{mode}_masked"].astype(np.int64)), f"masked tokens differ ({v}, {mode})"
            assert np.array_equal(mask.cpu().numpy(), g[f"v{v}_{mode}_mask"])
            for smooth, sum_splits in ((0.1, False), (0.0, True)):
                loss, d = MLMLoss(smooth, sum_splits)(logits.cuda(), tok, mask)
                got = np.array([float(d[k]) for k in ("mlm_loss", "correct_tokens", "masked_token_loss", "masked_correct_tokens")])
                want = g[f"v{v}_{mode}_loss_{smooth}_{int(sum_splits)}"]
                assert float(loss) == got[0]
                err = np.abs(got - want) / np.maximum(np.abs(want), 1e-6)
                print(f"MLMLoss V={v} {mode} smoothing={smooth} sum_splits={sum_splits}: {got} max rel err {err.max():.2e}")
                assert err.max() <= 2e-6                       # fp32 scalars of the reference vs double accumulation here


@pytest.mark.parametrize("v,b", [(32, 2), (512, 2), (64, 256)])
def test_mlm_loss_vs_oracle_other_shapes(v, b):
    """Other vocabularies and BASELINE config #2's batch (131 072 rows), against the fp64 oracle; determinism; the empty-mask case."""
    g = torch.Generator().manual_seed(v + b)
    logits = torch.randn((b, 256, 2, v), generator=g) * 3.0
    tok = torch.randint(0, v, (b, 256, 2), generator=g)
    mask = torch.rand((b, 256, 2), generator=g) < 0.4
    want = np.array(O.mlm_loss(logits, tok, mask, 0.1, False))
    lm = MLMLoss(0.1, False)
    _, d = lm(logits.cuda(), tok.cuda(), mask.cuda())
    got = np.array([float(d[k]) for k in ("mlm_loss", "correct_tokens", "masked_token_loss", "masked_correct_tokens")])
    assert (np.abs(got - want) <= 2e-6 * np.maximum(np.abs(want), 1e-3)).all(), (got, want)
    _, d2 = lm(logits.cuda(), tok.cuda(), mask.cuda())
    assert all(float(d[k]) == float(d2[k]) for k in d)                                   # fixed summation order
    _, d3 = lm(logits.cuda(), tok.cuda(), torch.zeros_like(mask).cuda())
    assert np.isnan(float(d3["masked_token_loss"])) and float(d3["mlm_loss"]) == float(d["mlm_loss"])   # mean over no rows, like torch


def test_training_forward_step_composition():
    """The forward half as train_maskbit.py:362-380 strings it together, on the CUDA path: encode -> split -> mask -> LFQBert forward
    -> MLMLoss, against the oracle on the CUDA path's own tokens and logits (the generator forward itself is pinned elsewhere)."""
    from maskbit_b200 import build_models, load_config
    cfg = load_config("maskbit_generator_12bit")
    tokenizer, gen = build_models(cfg, device="cuda")
    x = torch.rand((2, 3, 256, 256), generator=torch.Generator().manual_seed(3)).cuda()
    _, enc = tokenizer.encode(x)
    full = enc["min_encoding_indices"].reshape(2, -1)
    tok = split_factorized_tokens(full, 4096, 2)
    torch.manual_seed(5)
    masked, mask = get_mask_tokens(tok, cfg.model.mlm_model.mask_token, mode="arccos")
    assert torch.equal(masked[~mask], tok[~mask]) and (masked[mask] == 64).all()
    labels = torch.tensor([3, 900]).cuda()
    logits = gen(masked, labels, torch.tensor([False, True]).cuda())
    loss, d = MLMLoss(0.1, False)(logits, tok, mask)
    want = O.mlm_loss(logits.cpu(), tok.cpu(), mask.cpu(), 0.1, False)
    assert abs(float(loss) - want[0]) <= 2e-6 * want[0] and abs(float(d["masked_correct_tokens"]) - want[3]) <= 1e-6
    # random-init generator: the loss sits at log(V) = 4.16 within label-smoothing distance
    assert 3.9 < float(loss) < 4.5


def test_training_helpers_reject_cpu_tensors():
    with pytest.raises(maskbit_b200._lib.MaskbitError):
        split_factorized_tokens(torch.zeros((1, 4), dtype=torch.int64), 4096, 2)
    with pytest.raises(maskbit_b200._lib.MaskbitError):
        get_mask_tokens(torch.zeros((1, 4, 2), dtype=torch.int64), 64)
    with pytest.raises(ValueError):
        get_mask_tokens(torch.zeros((1, 4, 2), dtype=torch.int64).cuda(), 64, mode="root")
