"""llava-phi-3-mini language-model prefill on the C ABI (POL:456-463: `llava.generate(inputs_embeds=...)`, prefill part).

Packed variable-length batches (no padding): `inputs_embeds` [T, hidden] fp32 for all episodes of the rank, sequence
boundaries in `cu_seqlens`.  Returns the next-action logits of every sequence's last token.

Weights: Llama layout (q/k/v/o_proj, gate/up/down_proj) or Phi-3 layout (qkv_proj, gate_up_proj) state dicts are fused at
load into `w_qkv` [3*hidden, hidden] and a row-interleaved `w_gate_up` [2*ffn, hidden] (row 2j = gate_j, 2j+1 = up_j) so the
GEMM epilogue can apply SwiGLU in registers.  Residual stream / RMSNorm statistics fp32; GEMM operands 16-bit.
"""
import torch

from . import _lib as L
from . import ops


class LMWeights:
    @classmethod
    def from_state_dict(cls, sd, device="cuda", dtype=torch.float16, prefix="model.", lm_head_key="lm_head.weight"):
        w = cls()
        w.dtype = dtype
        c16 = lambda t: t.detach().to(device=device, dtype=torch.float32).to(dtype).contiguous()
        f32 = lambda t: t.detach().to(device=device, dtype=torch.float32).contiguous()
        w.embed = c16(sd[prefix + "embed_tokens.weight"])
        w.vocab, w.hidden = w.embed.shape
        w.norm = f32(sd[prefix + "norm.weight"])
        w.lm_head = c16(sd[lm_head_key])
        w.layers = []
        n_layers = len({k.split(".")[len(prefix.split("."))] for k in sd if k.startswith(prefix + "layers.")})
        for l in range(n_layers):
            p = f"{prefix}layers.{l}."
            if p + "self_attn.qkv_proj.weight" in sd:
                qkv = sd[p + "self_attn.qkv_proj.weight"]
                gu = sd[p + "mlp.gate_up_proj.weight"]
                gate, up = gu[: gu.shape[0] // 2], gu[gu.shape[0] // 2:]
            else:
                qkv = torch.cat([sd[p + f"self_attn.{n}_proj.weight"] for n in "qkv"], 0)
                gate, up = sd[p + "mlp.gate_proj.weight"], sd[p + "mlp.up_proj.weight"]
            inter = torch.stack([gate, up], dim=1).reshape(2 * gate.shape[0], gate.shape[1])
            w.layers.append({
                "rms1": f32(sd[p + "input_layernorm.weight"]), "w_qkv": c16(qkv), "w_o": c16(sd[p + "self_attn.o_proj.weight"]),
                "rms2": f32(sd[p + "post_attention_layernorm.weight"]), "w_gu": c16(inter), "w_down": c16(sd[p + "mlp.down_proj.weight"]),
            })
        w.ffn = w.layers[0]["w_down"].shape[1]
        return w


class LMEngine:
    def __init__(self, weights, n_heads=32, eps=1e-5, rope_theta=10000.0, max_tokens=1024, attention="auto"):
        L.require_device()
        self.w = weights
        self.H = n_heads
        self.Dh = weights.hidden // n_heads
        self.eps = eps
        self.attention = attention
        dev = weights.embed.device
        self.inv_freq = (1.0 / (rope_theta ** (torch.arange(0, self.Dh, 2, dtype=torch.float32) / self.Dh))).to(dev).contiguous()
        self._alloc(max_tokens)

    def _alloc(self, T):
        w, dev, dt = self.w, self.w.embed.device, self.w.dtype
        self.max_tokens = T
        self.A16 = torch.empty((T, w.hidden), device=dev, dtype=dt)
        self.qkv = torch.empty((T, 3 * w.hidden), device=dev, dtype=dt)
        self.att = torch.empty((T, w.hidden), device=dev, dtype=dt)
        self.h = torch.empty((T, w.ffn), device=dev, dtype=dt)

    def embed(self, ids, out):
        """embed_tokens (POL:439): ids int32 [T] -> out fp32 [T, hidden]."""
        ops.embed_gather(self.w.embed, ids, out)

    def prefill(self, X, cu_seqlens, positions, n_seq, max_len, last_rows):
        """X fp32 [T, hidden] (overwritten: it is the residual stream); cu_seqlens int32 [n_seq+1]; positions int32 [T];
        last_rows int32 [n_seq] (row of each sequence's last token).  Returns logits fp32 [n_seq, vocab]."""
        w = self.w
        ops.STAGE_TAG = "lm"
        T = X.shape[0]
        if T > self.max_tokens:
            self._alloc(T)
        A16, qkv, att, h = self.A16[:T], self.qkv[:T], self.att[:T], self.h[:T]
        tab = ops.rope_table(positions, self.inv_freq, self.Dh)  # cos/sin per token, shared by all layers
        for p in w.layers:
            ops.rmsnorm(X, p["rms1"], self.eps, out16=A16)
            ops.gemm(A16, p["w_qkv"], out=qkv)
            ops.rope_apply(qkv, tab, self.H, self.Dh)
            ops.attention(qkv, att, cu_seqlens, n_seq, max_len, self.H, self.Dh, causal=True, impl=self.attention)
            ops.gemm(att, p["w_o"], out=X, residual=X)
            ops.rmsnorm(X, p["rms2"], self.eps, out16=A16)
            ops.gemm(A16, p["w_gu"], out=h, act=L.ACT_SWIGLU)
            ops.gemm(h, p["w_down"], out=X, residual=X)
        last16 = torch.empty((n_seq, w.hidden), device=X.device, dtype=w.dtype)
        ops.rmsnorm(X, w.norm, self.eps, out16=last16, row_index=last_rows)
        logits = torch.empty((n_seq, w.vocab), device=X.device, dtype=torch.float32)
        ops.gemm(last16, w.lm_head, out=logits)
        return logits
