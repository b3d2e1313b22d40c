"""Select-path parity cases shared by tests/golden/make_golden.py (which runs the UNMODIFIED reference sample() on them) and the
tests (which regenerate the same inputs and compare the oracle / the CUDA kernel with the reference's recorded tokens).

The reference's select arithmetic (modeling/modules/sampling.py:90-131) cannot be called without a model, but sample() takes the
model as an argument: a STUB generator that returns seeded random logits lets the reference's own code make hundreds of thousands
of select decisions at every per-group vocabulary, without a 300 M-parameter forward and without storing the logits -- only the
reference's per-step tokens are committed (tests/golden/select_stub.npz), together with SHA-256 digests of the regenerated inputs
so that a torch build whose CPU generator differs fails loudly instead of silently testing other inputs.
"""
import hashlib

import numpy as np
import torch

# name -> case.  `kw` overrides the YAML's sampler kwargs (maskbit_b200.config.sampler_kwargs of `bits`).
CASES = {
    # the headline shape: 12-bit, V = 64, 16 x 16 steps x 512 slots = 131 072 decisions
    "v64_big": dict(bits=12, B=16, steps=16, seed=101, sigma_u=2.0, sigma_d=0.7, kw={}),
    # unguided branch (sampling.py:100-101) + temperature annealing (sampling.py:103)
    "v64_unguided_anneal": dict(bits=12, B=8, steps=8, seed=102, sigma_u=3.0, sigma_d=0.0,
                                kw=dict(guidance_scale=0.0, use_sampling_annealing=True, mask_schedule_strategy="cosine")),
    # peaked distributions: logit range of tens (what a trained checkpoint produces), exp underflow, p == 0 tokens, log(0) confidences
    "v64_peaked": dict(bits=12, B=4, steps=8, seed=103, sigma_u=9.0, sigma_d=3.0,
                       kw=dict(guidance_annealing="linear", mask_schedule_strategy="linear", softmax_temperature=0.7)),
    # BASELINE configs[2]: 14-bit, V = 128
    "v128_14bit": dict(bits=14, B=8, steps=12, seed=104, sigma_u=2.0, sigma_d=0.7, kw={}),
    "v32_10bit": dict(bits=10, B=4, steps=8, seed=105, sigma_u=2.0, sigma_d=1.0, kw={}),
    "v256_16bit": dict(bits=16, B=4, steps=8, seed=106, sigma_u=2.5, sigma_d=0.7, kw=dict(guidance_annealing="none", guidance_scale=3.0)),
    "v512_18bit": dict(bits=18, B=2, steps=8, seed=107, sigma_u=2.0, sigma_d=0.7, kw=dict(mask_schedule_strategy="root")),
}


def case_kwargs(case):
    from maskbit_b200.config import load_config, sampler_kwargs
    kw = dict(sampler_kwargs(load_config(f"maskbit_generator_{case['bits']}bit")), num_steps=case["steps"])
    kw.update(case["kw"])
    return kw


def stub_logits(case, step, guided):
    """Logits the stub generator returns at `step`: ([B,256,2,V] conditional, same unconditional or None)."""
    v = 2 ** (case["bits"] // 2)
    g = torch.Generator().manual_seed(case["seed"] * 1000 + step)
    lu = torch.randn((case["B"], 256, 2, v), generator=g) * case["sigma_u"]
    if not guided:
        return lu, None
    lc = lu + torch.randn((case["B"], 256, 2, v), generator=g) * case["sigma_d"]
    return lc, lu


class StubGenerator:
    """What sample() needs of `model` (sampling.py:55-58,84-101): .device, .eval() and the forward call."""

    def __init__(self, case, guided):
        self.case, self.guided, self.step, self.device = case, guided, 0, torch.device("cpu")

    def eval(self):
        return self

    def __call__(self, tokens, labels, drop):
        lc, lu = stub_logits(self.case, self.step, self.guided)
        self.step += 1
        assert tokens.shape[0] == (2 if self.guided else 1) * self.case["B"]
        return torch.cat([lc, lu], 0) if self.guided else lc


class StubTokenizer:
    def eval(self):
        return self

    def decode_tokens(self, tokens):
        return torch.zeros((tokens.shape[0], 3, 1, 1))


NOISE_SEED = 4321


def replay_noise(case):
    """The global-generator draws sample() makes on CPU for this case (per step: B*512*V exponentials, then the Gumbel sample)."""
    v = 2 ** (case["bits"] // 2)
    torch.manual_seed(NOISE_SEED)
    gumbel = torch.distributions.Gumbel(loc=0.0, scale=1.0)
    qs, gs = [], []
    for _ in range(case["steps"]):
        qs.append(torch.empty(case["B"] * 512, v).exponential_(1))
        gs.append(gumbel.sample((case["B"], 256, 2)))
    return qs, gs


def digest(tensors):
    h = hashlib.sha256()
    for t in tensors:
        if t is not None:
            h.update(np.ascontiguousarray(t.numpy()).tobytes())
    return h.hexdigest()


def case_inputs(case):
    """(kwargs, guided, [(lc, lu)] per step, q per step, gumbel per step, input digest)."""
    kw = case_kwargs(case)
    guided = kw["guidance_scale"] != 0.0
    logits = [stub_logits(case, i, guided) for i in range(case["steps"])]
    qs, gs = replay_noise(case)
    dg = digest([t for pair in logits for t in pair] + qs + gs)
    return kw, guided, logits, qs, gs, dg
