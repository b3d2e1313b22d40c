"""Checkpoint / device plumbing shared by the two model mirrors.

Mirrors modeling/modules/base_model.py (load_pretrained :87-142, save_pretrained :48-85, device/dtype :145-157):
same file layout (``pytorch_model.bin`` = torch.save(state_dict)), same ``rename_keys`` prefix semantics, strict
loading, eval mode.  The weights themselves live inside libmaskbit_b200 (bf16 / split-bf16 packed); this class
keeps the fp32 CPU state_dict only so that ``state_dict()`` / ``save_pretrained`` round-trip.
"""
import ctypes
import os
from collections import OrderedDict

import torch

from . import _lib


class EngineModel:
    _model_id = None  # MB_GENERATOR or MB_TOKENIZER

    def __init__(self):
        self._sd = None
        self._device = torch.device("cpu")
        self._handle = None
        self.training = False

    # ---- to be provided by subclasses
    def _mb_config(self):
        raise NotImplementedError

    def _expected_spec(self):
        """list of (name, shape) the state_dict must contain"""
        raise NotImplementedError

    def _default_state_dict(self):
        raise NotImplementedError

    # ---- torch.nn.Module-like surface used by the reference's callers
    def eval(self):
        self.training = False
        return self

    def train(self, mode=True):
        if mode:
            raise NotImplementedError("maskbit_b200 implements the inference path only")
        return self

    def requires_grad_(self, requires_grad=False):
        if requires_grad:
            raise NotImplementedError("maskbit_b200 implements the inference path only")
        return self

    @property
    def device(self):
        return self._device

    @property
    def dtype(self):
        return torch.float32

    def state_dict(self):
        self._ensure_weights()
        return OrderedDict(self._sd)

    def load_state_dict(self, state_dict, strict=True):
        spec = self._expected_spec()
        names = [n for n, _ in spec]
        nameset = set(names)
        missing = [n for n in names if n not in state_dict]
        unexpected = [k for k in state_dict if k not in nameset]
        if strict and (missing or unexpected):
            raise RuntimeError(f"Error(s) in loading state_dict for {type(self).__name__}: "
                               f"Missing key(s): {missing[:8]} Unexpected key(s): {unexpected[:8]}")
        for n, shape in spec:
            if n in state_dict and tuple(state_dict[n].shape) != tuple(shape):
                raise RuntimeError(f"size mismatch for {n}: checkpoint {tuple(state_dict[n].shape)} vs model {tuple(shape)}")
        if missing:
            base = self._default_state_dict()
            base.update({k: v for k, v in state_dict.items() if k in nameset})
            state_dict = base
        self._sd = OrderedDict((n, state_dict[n].detach().to("cpu")) for n in names)
        self._release()
        if self._device.type == "cuda":
            self._upload()
        return self

    @staticmethod
    def _checkpoint_file(path):
        """A checkpoint file itself, or `pytorch_model.bin` inside a checkpoint directory (base_model.py:108-117)."""
        candidate = os.path.join(path, "pytorch_model.bin") if os.path.isdir(path) else path
        if not os.path.isfile(candidate):
            raise ValueError(f"{candidate} does not exist")
        return candidate

    def load_pretrained(self, pretrained_model_path, strict_loading=True, torch_dtype=None, rename_keys=None):
        """Same contract as the reference's BaseModel.load_pretrained (base_model.py:87-142): load a state dict from a file or a
        checkpoint directory, optionally rename keys (`rename_keys` maps an old key prefix to its replacement; the first matching
        prefix wins and, like the reference's str.replace, every occurrence inside the key is replaced), load strictly unless told
        otherwise, end in eval mode."""
        loaded = torch.load(self._checkpoint_file(pretrained_model_path), map_location="cpu")

        def renamed(key):
            hit = next((old for old in (rename_keys or {}) if key.startswith(old)), None)
            return key if hit is None else key.replace(hit, rename_keys[hit])

        self.load_state_dict({renamed(k): v for k, v in loaded.items()}, strict=strict_loading)
        if torch_dtype is not None and not isinstance(torch_dtype, torch.dtype):
            raise ValueError(f"{torch_dtype} needs to be of type `torch.dtype`, e.g. `torch.float16`, but is {type(torch_dtype)}.")
        if torch_dtype not in (None, torch.float32):
            raise NotImplementedError("the engine chooses its own compute precision; only torch.float32 checkpoints are accepted")
        self.eval()

    def save_pretrained(self, save_directory, save_function=None, state_dict=None):
        """base_model.py:48-85."""
        if os.path.isfile(save_directory):
            print(f"Provided path ({save_directory}) should be a directory, not a file")
            return
        save_function = save_function or torch.save
        os.makedirs(save_directory, exist_ok=True)
        save_function(state_dict if state_dict is not None else self.state_dict(),
                      os.path.join(save_directory, "pytorch_model.bin"))

    def to(self, device=None, *args, **kwargs):
        if device is None or isinstance(device, torch.dtype):
            return self
        device = torch.device(device)
        if device.type == "cuda" and device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        if device != self._device:
            self._release()
            self._device = device
            if device.type == "cuda" and self._sd is not None:
                self._upload()
        return self

    def cuda(self, device=None):
        return self.to(torch.device("cuda", device if device is not None else torch.cuda.current_device()))

    def cpu(self):
        return self.to("cpu")

    # ---- engine handle management
    def _ensure_weights(self):
        if self._sd is None:
            self._sd = self._default_state_dict()

    def _release(self):
        if self._handle is not None:
            _lib.lib().mb_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def _upload(self):
        self._ensure_weights()
        L = _lib.lib()
        cfg = self._mb_config()
        with torch.cuda.device(self._device):
            h = ctypes.c_void_p()
            _lib.check(L.mb_create(ctypes.byref(cfg), ctypes.byref(h)))
            try:
                for name, t in self._sd.items():
                    t32 = t.to(torch.float32).contiguous()
                    shape = (ctypes.c_int64 * max(1, t32.dim()))(*t32.shape)
                    _lib.check(L.mb_set_tensor(h, self._model_id, name.encode(), ctypes.c_void_p(t32.data_ptr()), shape,
                                               t32.dim(), 0))
                _lib.check(L.mb_finalize(h, self._model_id))
            except Exception:
                L.mb_destroy(h)
                raise
        self._handle = h

    def _engine(self):
        """The live handle; the product path requires a CUDA device and the built extension."""
        if self._device.type != "cuda":
            raise _lib.MaskbitError(f"{type(self).__name__} is on {self._device}: maskbit_b200 runs on a B200 (sm_100a) only -- "
                                    "call .to('cuda'); there is no CPU fallback")
        if self._handle is None:
            self._upload()
        return self._handle

    def launch_count(self):
        return int(_lib.lib().mb_launch_count(self._handle)) if self._handle is not None else 0

    def profile_enable(self, on=True):
        """Start (or stop) per-kernel-class CUDA-event timing on this model's handle (mb_profile_enable)."""
        _lib.check(_lib.lib().mb_profile_enable(self._engine(), int(bool(on))))

    def profile_read(self):
        """{kernel class: (milliseconds, launches)} since the last read; synchronises the device."""
        n = len(_lib.PROF_KINDS)
        ms = (ctypes.c_double * n)()
        cnt = (ctypes.c_int64 * n)()
        _lib.check(_lib.lib().mb_profile_read(self._engine(), ms, cnt, n))
        return {k: (ms[i], int(cnt[i])) for i, k in enumerate(_lib.PROF_KINDS) if cnt[i]}
