"""CLIP ViT-L/14@336 patch encoder on the C ABI (tcgen05 GEMMs + fused norm / attention kernels).

Drop-in for the reference's `CLIPEncoder` (ENC:245-284) and, with `n_layers_run=23, ln_post=False`, for the LLaVA
vision tower's `hidden_states[-2]` (POL:441-452).  Weight layouts accepted: OpenAI CLIP `visual.*` state dict
(CLIPM:203-238) and HF `CLIPVisionModel` (`vision_model.*`).

Precision: weights / GEMM operands fp16 (the reference converts CLIP to fp16, CLIPM:389-410); residual stream,
LayerNorm and softmax statistics fp32; fp32 accumulation.
"""
import torch

from . import _lib as L
from . import ops


class ViTWeights:
    """Device-resident ViT parameters in engine layout."""

    def __init__(self):
        self.layers = []

    @staticmethod
    def _w16(t, device, dtype):
        return t.detach().to(device=device, dtype=torch.float32).to(dtype).contiguous()

    @staticmethod
    def _f32(t, device):
        return t.detach().to(device=device, dtype=torch.float32).contiguous()

    @classmethod
    def from_openai_state_dict(cls, sd, device="cuda", dtype=torch.float16, prefix=""):
        """`sd` uses the OpenAI names (conv1.weight, transformer.resblocks.N.*, ln_post.*, proj); `prefix` e.g. 'visual.'."""
        g = lambda k: sd[prefix + k]
        w = cls()
        w.dtype = dtype
        conv = g("conv1.weight")
        w.width, _, w.patch, _ = conv.shape
        k = 3 * w.patch * w.patch
        w.kpad = (k + 7) // 8 * 8
        cw = torch.zeros((w.width, w.kpad), dtype=torch.float32)
        cw[:, :k] = conv.detach().float().reshape(w.width, k).cpu()
        w.conv_w = cls._w16(cw, device, dtype)
        w.cls = cls._f32(g("class_embedding"), device)
        w.pos = cls._f32(g("positional_embedding"), device)
        w.tokens = w.pos.shape[0]
        w.ln_pre = (cls._f32(g("ln_pre.weight"), device), cls._f32(g("ln_pre.bias"), device))
        n_layers = len({k_[len(prefix):].split(".")[2] for k_ in sd if k_.startswith(prefix + "transformer.resblocks.")})
        for l in range(n_layers):
            p = f"transformer.resblocks.{l}."
            w.layers.append({
                "ln1": (cls._f32(g(p + "ln_1.weight"), device), cls._f32(g(p + "ln_1.bias"), device)),
                "w_qkv": cls._w16(g(p + "attn.in_proj_weight"), device, dtype), "b_qkv": cls._f32(g(p + "attn.in_proj_bias"), device),
                "w_o": cls._w16(g(p + "attn.out_proj.weight"), device, dtype), "b_o": cls._f32(g(p + "attn.out_proj.bias"), device),
                "ln2": (cls._f32(g(p + "ln_2.weight"), device), cls._f32(g(p + "ln_2.bias"), device)),
                "w_fc": cls._w16(g(p + "mlp.c_fc.weight"), device, dtype), "b_fc": cls._f32(g(p + "mlp.c_fc.bias"), device),
                "w_pr": cls._w16(g(p + "mlp.c_proj.weight"), device, dtype), "b_pr": cls._f32(g(p + "mlp.c_proj.bias"), device),
            })
        if prefix + "ln_post.weight" in sd:
            w.ln_post = (cls._f32(g("ln_post.weight"), device), cls._f32(g("ln_post.bias"), device))
            w.proj = cls._w16(g("proj").detach().float().t(), device, dtype)  # [out, width]
            w.out_dim = w.proj.shape[0]
        else:
            w.ln_post, w.proj, w.out_dim = None, None, w.width
        return w

    @classmethod
    def from_hf_clip_state_dict(cls, sd, device="cuda", dtype=torch.float16, prefix="vision_model."):
        """HF CLIPVisionModel names -> OpenAI names, then `from_openai_state_dict` (q/k/v projections are fused)."""
        g = lambda k: sd[prefix + k]
        o = {"conv1.weight": g("embeddings.patch_embedding.weight"), "class_embedding": g("embeddings.class_embedding"),
             "positional_embedding": g("embeddings.position_embedding.weight"),
             "ln_pre.weight": g("pre_layrnorm.weight"), "ln_pre.bias": g("pre_layrnorm.bias")}
        n_layers = len({k_.split(".")[len(prefix.split(".")) + 1] for k_ in sd if k_.startswith(prefix + "encoder.layers.")})
        for l in range(n_layers):
            s, d = f"encoder.layers.{l}.", f"transformer.resblocks.{l}."
            o[d + "ln_1.weight"], o[d + "ln_1.bias"] = g(s + "layer_norm1.weight"), g(s + "layer_norm1.bias")
            o[d + "ln_2.weight"], o[d + "ln_2.bias"] = g(s + "layer_norm2.weight"), g(s + "layer_norm2.bias")
            o[d + "attn.in_proj_weight"] = torch.cat([g(s + f"self_attn.{x}_proj.weight") for x in "qkv"], 0)
            o[d + "attn.in_proj_bias"] = torch.cat([g(s + f"self_attn.{x}_proj.bias") for x in "qkv"], 0)
            o[d + "attn.out_proj.weight"], o[d + "attn.out_proj.bias"] = g(s + "self_attn.out_proj.weight"), g(s + "self_attn.out_proj.bias")
            o[d + "mlp.c_fc.weight"], o[d + "mlp.c_fc.bias"] = g(s + "mlp.fc1.weight"), g(s + "mlp.fc1.bias")
            o[d + "mlp.c_proj.weight"], o[d + "mlp.c_proj.bias"] = g(s + "mlp.fc2.weight"), g(s + "mlp.fc2.bias")
        return cls.from_openai_state_dict(o, device, dtype)


class ViTEngine:
    """Runs the tower for up to `max_images` images per call; all scratch is preallocated (CUDA-graph friendly)."""

    def __init__(self, weights, n_head=16, resolution=336, max_images=12, attention="auto", tag="vit"):
        L.require_device()
        self.tag = tag  # stage name in ops.STAGE_PROFILE
        self.w = weights
        self.H = n_head
        self.R = resolution
        self.g2 = (resolution // weights.patch) ** 2
        assert self.g2 + 1 == weights.tokens
        self.attention = attention
        self._alloc(max_images)

    def _alloc(self, n):
        w, dev, dt = self.w, self.w.conv_w.device, self.w.dtype
        T = n * w.tokens
        self.max_images = n
        self.conv = torch.empty((n * self.g2, w.width), device=dev, dtype=torch.float32)
        self.X = torch.empty((T, w.width), device=dev, dtype=torch.float32)
        self.A16 = torch.empty((T, w.width), device=dev, dtype=dt)
        self.qkv = torch.empty((T, 3 * w.width), device=dev, dtype=dt)
        self.att = torch.empty((T, w.width), device=dev, dtype=dt)
        self.h = torch.empty((T, 4 * w.width), device=dev, dtype=dt)
        self.out = torch.empty((T, w.out_dim), device=dev, dtype=dt)
        self.cu = (torch.arange(n + 1, device=dev, dtype=torch.int32) * w.tokens).contiguous()
        self.cols = torch.empty((n * self.g2, w.kpad), device=dev, dtype=dt)
        self._c = None

    def _c_model(self):
        """(d3d_vit_model, d3d_vit_scratch) for the one-call forward (csrc/forward_host.cu); rebuilt when the scratch was re-allocated."""
        if self._c is None:
            w = self.w
            arr = (L.ViTLayer * len(w.layers))()
            for i, p in enumerate(w.layers):
                arr[i] = L.ViTLayer(p["ln1"][0].data_ptr(), p["ln1"][1].data_ptr(), p["w_qkv"].data_ptr(), p["b_qkv"].data_ptr(), p["w_o"].data_ptr(),
                                    p["b_o"].data_ptr(), p["ln2"][0].data_ptr(), p["ln2"][1].data_ptr(), p["w_fc"].data_ptr(), p["b_fc"].data_ptr(),
                                    p["w_pr"].data_ptr(), p["b_pr"].data_ptr())
            m = L.ViTModel(len(w.layers), w.width, self.H, w.patch, w.tokens, w.kpad, self.R, w.out_dim, L.kind_of(w.dtype), w.conv_w.data_ptr(),
                           w.cls.data_ptr(), w.pos.data_ptr(), w.ln_pre[0].data_ptr(), w.ln_pre[1].data_ptr(), arr,
                           w.ln_post[0].data_ptr() if w.ln_post is not None else None, w.ln_post[1].data_ptr() if w.ln_post is not None else None,
                           w.proj.data_ptr() if w.proj is not None else None)
            sc = L.ViTScratch(self.cols.data_ptr(), self.conv.data_ptr(), self.X.data_ptr(), self.A16.data_ptr(), self.qkv.data_ptr(),
                              self.att.data_ptr(), self.h.data_ptr(), self.out.data_ptr(), self.cu.data_ptr())
            self._c = (m, sc, arr)
        return self._c[0], self._c[1]

    def forward(self, img_u8, n_layers_run=None, ln_post_on_patches=True, project=True):
        """img_u8 [N,H,W,3] uint8 on device.  Returns (cls [N,out], patch [N,g2,out]) 16-bit views into `self.out`
        (valid until the next call), or with project=False the fp32 hidden state [N,tokens,width] after `n_layers_run` blocks."""
        w = self.w
        ops.STAGE_TAG = self.tag
        N = img_u8.shape[0]
        if N > self.max_images:
            self._alloc(N)
        T = N * w.tokens
        run = len(w.layers) if n_layers_run is None else n_layers_run
        if ops.STAGE_PROFILE is None and self.attention == "auto" and (ln_post_on_patches or not project):
            # the whole forward behind ONE C call (d3d_vit_forward); the per-kernel loop below remains for the per-stage profile, explicit
            # attention implementations and the Pretrain (ln_post-less, Q8) projection
            import ctypes
            m, sc = self._c_model()
            L.check(L.lib().d3d_vit_forward(ctypes.addressof(m), L.ptr(img_u8), N, img_u8.shape[1], img_u8.shape[2], run, int(bool(project)),
                                            ctypes.addressof(sc), L.stream_ptr()))
            if not project:
                return self.X[:T].view(N, w.tokens, w.width)
            o = self.out[:T].view(N, w.tokens, w.out_dim)
            return o[:, 0], o[:, 1:]
        X, A16, qkv, att, h = self.X[:T], self.A16[:T], self.qkv[:T], self.att[:T], self.h[:T]
        cols = ops.preprocess_im2col(img_u8, self.R, w.patch, w.dtype)
        ops.gemm(cols, w.conv_w, out=self.conv[:N * self.g2])
        ops.vit_embed_ln(self.conv, w.cls, w.pos, w.ln_pre[0], w.ln_pre[1], 1e-5, N, w.tokens, X)
        Dh = w.width // self.H
        for l in range(run):
            p = w.layers[l]
            ops.layernorm(X, p["ln1"][0], p["ln1"][1], 1e-5, out16=A16)
            ops.gemm(A16, p["w_qkv"], out=qkv, bias=p["b_qkv"])
            ops.attention(qkv, att, self.cu, N, w.tokens, self.H, Dh, causal=False, uniform_len=w.tokens, impl=self.attention)
            ops.gemm(att, p["w_o"], out=X, bias=p["b_o"], residual=X)
            ops.layernorm(X, p["ln2"][0], p["ln2"][1], 1e-5, out16=A16)
            ops.gemm(A16, p["w_fc"], out=h, bias=p["b_fc"], act=L.ACT_QUICK_GELU)
            ops.gemm(h, p["w_pr"], out=X, bias=p["b_pr"], residual=X)
        if not project:
            return X.view(N, w.tokens, w.width)
        if ln_post_on_patches:
            ops.layernorm(X, w.ln_post[0], w.ln_post[1], 1e-5, out16=A16)
        else:  # Pretrain variant (Q8): ln_post on CLS only, raw patch tokens are projected
            ops.cast16(X, A16)
            cls_rows = (torch.arange(N, device=X.device, dtype=torch.int32) * w.tokens).contiguous()
            tmp = torch.empty((N, w.width), device=X.device, dtype=w.dtype)
            ops.layernorm(X, w.ln_post[0], w.ln_post[1], 1e-5, out16=tmp, row_index=cls_rows)
            A16.view(N, w.tokens, w.width)[:, 0].copy_(tmp)
        out = self.out[:T]
        ops.gemm(A16, w.proj, out=out)
        o = out.view(N, w.tokens, w.out_dim)
        return o[:, 0], o[:, 1:]
