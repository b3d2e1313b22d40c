"""ConvVQModel -- drop-in mirror of modeling/conv_vqgan.py:39-132 (decode_tokens / decode / encode / forward); the
convolutions run in libmaskbit_b200's kernels."""
import ctypes

import torch

from . import _lib
from .base_model import EngineModel
from .weights import conv_vq_spec, synthetic_conv_vq_state_dict


class ConvVQModel(EngineModel):
    _model_id = _lib.MB_TOKENIZER

    def __init__(self, config, legacy: bool = False, finetune_decoder: bool = False):
        super().__init__()
        if legacy:
            raise NotImplementedError("legacy=True (ConvDecoderLegacy) is outside the MaskBit sampling path (eval_maskbit.py:26 uses legacy=False)")
        if config.quantizer_type != "lookup-free":
            raise NotImplementedError("only the lookup-free quantizer is on the MaskBit sampling path")
        if not config.get("sample_with_conv", True):
            raise NotImplementedError("sample_with_conv=False (pooling / plain interpolation stages, autoencoder.py:160-183,216-227) is not built: "
                                      "every shipped tokenizer config uses strided / upsample convolutions")
        self.config = config
        self.finetune_decoder = finetune_decoder
        self.token_size = int(config.token_size)

    def _arch(self):
        c = self.config
        return dict(token_size=self.token_size, num_channels=c.num_channels, hidden_channels=c.hidden_channels,
                    channel_mult=tuple(c.channel_mult), num_resolutions=c.num_resolutions,
                    num_res_blocks=c.get("num_res_blocks_decoder", c.num_res_blocks),      # ConvDecoder (autoencoder.py:371)
                    num_res_blocks_encoder=c.num_res_blocks)                               # ConvEncoder (autoencoder.py:245)

    def _expected_spec(self):
        return [(n, s) for n, s, _ in conv_vq_spec(**self._arch())]

    def _default_state_dict(self):
        return synthetic_conv_vq_state_dict(seed=0, **self._arch())

    def _mb_config(self):
        a = self._arch()
        c = _lib.MBConfig()
        # generator-side fields are unused by a tokenizer handle but must pass validation
        c.hidden_dim, c.depth, c.heads, c.mlp_dim = 1024, 1, 16, 4096
        c.token_bits, c.codebook_splits, c.nclass, c.seq_len = self.token_size, 1 if self.token_size % 2 else 2, 1000, 256
        c.use_prenorm = 0
        c.dec_hidden_channels = a["hidden_channels"]
        for i, v in enumerate(a["channel_mult"]):
            c.dec_channel_mult[i] = v
        c.dec_num_resolutions, c.dec_num_res_blocks, c.num_channels = a["num_resolutions"], a["num_res_blocks"], a["num_channels"]
        c.enc_num_res_blocks = a["num_res_blocks_encoder"]
        return c

    @property
    def image_size(self):
        return 16 * 2 ** (self.config.num_resolutions - 1)

    @torch.no_grad()
    def decode_tokens(self, tokens):
        """conv_vqgan.py:98-112: tokens [B, 256] (any int / float dtype, `.long()`-ed like lookup_free.py:108)
        -> fp32 [B, 3, H, W], unclamped."""
        h = self._engine()
        dev = self._device
        tok = tokens.to(device=dev).long().contiguous()
        if tok.dim() != 2 or tok.shape[1] != 256:
            raise ValueError(f"expected tokens [B,256], got {tuple(tok.shape)}")
        b = tok.shape[0]
        with torch.cuda.device(dev):
            img = torch.empty((b, 3, self.image_size, self.image_size), dtype=torch.float32, device=dev)
            if b > 0:
                _lib.check(_lib.lib().mb_decode_tokens(h, ctypes.c_void_p(tok.data_ptr()), b, ctypes.c_void_p(img.data_ptr()),
                                                       _lib.current_stream()))
        return img

    @torch.no_grad()
    def decode(self, z_quantized):
        """conv_vqgan.py:86-96 for lookup-free latents: z [B, bits, 16, 16] with entries in {-1, +1}."""
        z = z_quantized.to(self._device)
        if not bool(((z == 1) | (z == -1)).all()):
            raise NotImplementedError("decode() accepts LFQ latents (entries +-1) only; conv_in is fused with the bit unpack")
        b2i = (2 ** torch.arange(z.shape[1], device=z.device)).view(1, -1, 1, 1)
        tokens = ((z > 0).long() * b2i).sum(1).reshape(z.shape[0], -1)
        return self.decode_tokens(tokens)

    @torch.no_grad()
    def postprocess_uint8(self, images):
        """scripts/eval_maskbit.py:134-135: clamp(0,1)*255 -> permute(0,2,3,1) -> uint8, on the device."""
        h = self._engine()
        b = images.shape[0]
        with torch.cuda.device(self._device):
            out = torch.empty((b, self.image_size, self.image_size, 3), dtype=torch.uint8, device=self._device)
            _lib.check(_lib.lib().mb_postprocess_u8(h, ctypes.c_void_p(images.contiguous().data_ptr()), b,
                                                    ctypes.c_void_p(out.data_ptr()), _lib.current_stream()))
        return out

    @torch.no_grad()
    def tokenize(self, x, return_latents=False):
        """Images fp32 [B,3,H,W] -> LFQ tokens int64 [B,16,16] (``min_encoding_indices`` of lookup_free.py:60) and, when
        asked, the encoder latents z fp32 [B,bits,16,16] before the sign."""
        h = self._engine()
        dev = self._device
        x = x.to(device=dev, dtype=torch.float32).contiguous()
        if x.dim() != 4 or x.shape[1] != 3 or x.shape[2] != self.image_size or x.shape[3] != self.image_size:
            raise ValueError(f"expected images [B,3,{self.image_size},{self.image_size}], got {tuple(x.shape)}")
        b = x.shape[0]
        with torch.cuda.device(dev):
            idx = torch.empty((b, 16, 16), dtype=torch.int64, device=dev)
            z = torch.empty((b, self.token_size, 16, 16), dtype=torch.float32, device=dev) if return_latents else None
            if b > 0:
                _lib.check(_lib.lib().mb_encode(h, ctypes.c_void_p(x.data_ptr()), b, ctypes.c_void_p(z.data_ptr()) if z is not None else None,
                                                ctypes.c_void_p(idx.data_ptr()), _lib.current_stream()))
        return (idx, z) if return_latents else idx

    @torch.no_grad()
    def encode(self, x):
        """conv_vqgan.py:71-84 in eval mode: (z_quantized [B,bits,16,16] with entries +-1, result_dict).  The reference returns
        z + (sign(z) - z), i.e. +-1 up to fp32 rounding (lookup_free.py:80); here the entries are exactly +-1.  Losses are
        training quantities: commitment_loss / quantizer_loss are evaluated from the latents, the entropy terms are 0 in
        eval mode exactly as in the reference (lookup_free.py:64-74)."""
        idx, z = self.tokenize(x, return_latents=True)
        b2i = (2 ** torch.arange(self.token_size, device=idx.device)).view(1, -1, 1, 1)
        zq = ((idx.unsqueeze(1) & b2i) != 0).float() * 2.0 - 1.0
        zero = torch.zeros((), device=idx.device)
        commitment = float(self.config.commitment_cost) * torch.mean((zq - z) ** 2)
        result = dict(quantizer_loss=commitment, commitment_loss=commitment, entropy_loss=zero, per_sample_entropy=zero,
                      avg_entropy=zero, min_encoding_indices=idx)
        return zq, result

    @torch.no_grad()
    def forward(self, x):
        """conv_vqgan.py:114-132: (reconstruction fp32 [B,3,H,W], result_dict)."""
        zq, result = self.encode(x)
        return self.decode_tokens(result["min_encoding_indices"].reshape(zq.shape[0], -1)), result

    __call__ = forward
