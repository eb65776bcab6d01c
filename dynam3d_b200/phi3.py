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
        self.trim_last_layer = True  # final layer: o_proj / MLP on the last-token rows only (d3d_phi3_prefill)
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
        self.rope_tab = torch.empty((T, self.Dh), device=dev, dtype=torch.float32)
        self.last16 = torch.empty((64, w.hidden), device=dev, dtype=dt)
        self.att_last = torch.empty((64, w.hidden), device=dev, dtype=dt)
        self.x_last = torch.empty((64, w.hidden), device=dev, dtype=torch.float32)

    def embed(self, ids, out):
        """embed_tokens (POL:439): ids int32 [T] -> out fp32 [T, hidden]."""
        ops.embed_gather(self.w.embed, ids, out)

    def prefill(self, X, cu_seqlens, positions, n_seq, max_len, last_rows, kv_rows=0):
        """X fp32 [T, hidden] (overwritten: it is the residual stream); cu_seqlens int32 [n_seq+1]; positions int32 [T];
        last_rows int32 [n_seq] (row of each sequence's last token).  Returns logits fp32 [n_seq, vocab].
        kv_rows > 0 keeps every layer's packed QKV matrix (K rotated) in `self.kv` [layers, T + kv_rows, 3*hidden]: the KV cache of `generate`."""
        w = self.w
        ops.STAGE_TAG = "lm"
        T = X.shape[0]
        if T > self.max_tokens:
            self._alloc(T)
        A16, qkv, att, h = self.A16[:T], self.qkv[:T], self.att[:T], self.h[:T]
        if kv_rows:
            kv = getattr(self, "kv", None)
            if kv is None or kv.shape[1] < T + kv_rows:
                self.kv = kv = torch.empty((len(w.layers), T + kv_rows, 3 * w.hidden), device=X.device, dtype=w.dtype)
        if ops.STAGE_PROFILE is None and self.attention == "auto" and n_seq <= 64 and X.is_contiguous():
            # the whole prefill behind ONE C call (d3d_phi3_prefill, csrc/forward_host.cu); the per-kernel loop below remains for the per-stage
            # profile.  The final layer's o_proj / MLP run on the last-token rows only (the logits are the only output).
            import ctypes
            m = self._c_model()
            sc = L.LMScratch(self.A16.data_ptr(), self.qkv.data_ptr(), self.att.data_ptr(), self.h.data_ptr(), self.rope_tab.data_ptr(),
                             self.last16.data_ptr(), self.att_last.data_ptr(), self.x_last.data_ptr())
            ptrs = (ctypes.c_void_p * len(w.layers))(*[self.kv[l].data_ptr() for l in range(len(w.layers))]) if kv_rows else None
            logits = torch.empty((n_seq, w.vocab), device=X.device, dtype=torch.float32)
            L.check(L.lib().d3d_phi3_prefill(ctypes.addressof(m), L.ptr(X), T, L.ptr(cu_seqlens), L.ptr(positions), n_seq, int(max_len), L.ptr(last_rows),
                                             L.ptr(self.inv_freq), ctypes.cast(ptrs, ctypes.c_void_p) if kv_rows else None,
                                             self.kv.stride(1) if kv_rows else 0, ctypes.addressof(sc), int(self.trim_last_layer), L.ptr(logits),
                                             L.stream_ptr()))
            return logits
        tab = ops.rope_table(positions, self.inv_freq, self.Dh)  # cos/sin per token, shared by all layers
        for li, p in enumerate(w.layers):
            if kv_rows:
                qkv = self.kv[li, :T]
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

    # ------------------------------------------------------------------ chunked prefill over a strided KV cache (d3d_phi3_prefill_chunk)
    def chunk_cache(self, n_seq, stride):
        """Per-layer packed-QKV cache [layers, n_seq * stride, 3*hidden] (sequence s at rows s*stride + position) and the attention scratch.
        Zero-initialised ONCE: rows past a sequence's end are read (masked) by the attention tiles, so they must stay finite."""
        w = self.w
        kc = getattr(self, "kvc", None)
        if kc is None or kc.shape[1] < n_seq * stride or getattr(self, "kvc_stride", 0) != stride:
            dev = w.embed.device
            self.kvc = torch.zeros((len(w.layers), n_seq * stride, 3 * w.hidden), device=dev, dtype=w.dtype)
            self.attc = torch.zeros((n_seq * stride, w.hidden), device=dev, dtype=w.dtype)
            self.kvc_stride = stride
            self._kvc_ptrs = None
        return self.kvc

    def prefill_chunk(self, X, seq_start, seq_len, rows, positions, n_seq, max_len, q_tile_begin, q_tile_end, last_rows=None):
        """One chunk of a prefill (see include/dynam3d_b200.h: d3d_phi3_prefill_chunk).  X fp32 [T, hidden] compact rows (overwritten);
        seq_start / seq_len int32 [n_seq], rows / positions int32 [T] (device).  Returns logits [n_seq, vocab] when last_rows is given."""
        import ctypes
        w = self.w
        T = X.shape[0]
        if T > self.max_tokens:
            self._alloc(T)
        assert X.is_contiguous() and self.kvc is not None
        if getattr(self, "_kvc_ptrs", None) is None:
            self._kvc_ptrs = (ctypes.c_void_p * len(w.layers))(*[self.kvc[l].data_ptr() for l in range(len(w.layers))])
        m = self._c_model()
        sc = L.LMScratch(self.A16.data_ptr(), self.qkv.data_ptr(), self.att.data_ptr(), self.h.data_ptr(), self.rope_tab.data_ptr(),
                         self.last16.data_ptr(), self.att_last.data_ptr(), self.x_last.data_ptr())
        ch = L.LMChunk(seq_start.data_ptr(), seq_len.data_ptr(), rows.data_ptr(), positions.data_ptr(), int(q_tile_begin), int(q_tile_end))
        logits = torch.empty((n_seq, w.vocab), device=X.device, dtype=torch.float32) if last_rows is not None else None
        L.check(L.lib().d3d_phi3_prefill_chunk(ctypes.addressof(m), L.ptr(X), T, n_seq, int(max_len), ctypes.addressof(ch),
                                               ctypes.cast(self._kvc_ptrs, ctypes.c_void_p), self.kvc.stride(1), self.kvc.shape[1], L.ptr(self.attc),
                                               L.ptr(self.inv_freq), ctypes.addressof(sc), L.ptr(last_rows), int(self.trim_last_layer), L.ptr(logits),
                                               L.stream_ptr()))
        return logits

    # ------------------------------------------------------------------ greedy decode with the KV cache (POL:463-469)
    def _c_model(self):
        if getattr(self, "_cm", None) is None:
            w = self.w
            arr = (L.LMLayer * len(w.layers))()
            for i, p in enumerate(w.layers):
                arr[i] = L.LMLayer(p["rms1"].data_ptr(), p["w_qkv"].data_ptr(), p["w_o"].data_ptr(), p["rms2"].data_ptr(), p["w_gu"].data_ptr(),
                                   p["w_down"].data_ptr())
            m = L.LMModel(len(w.layers), w.hidden, self.H, self.Dh, w.ffn, w.vocab, L.kind_of(w.dtype), self.eps, arr, w.norm.data_ptr(),
                          w.lm_head.data_ptr(), w.embed.data_ptr())
            self._cm = (m, arr)
        return self._cm[0]

    def generate(self, X, cu_seqlens, positions, n_seq, max_len, last_rows, max_new_tokens=20, eos_ids=(), step_logits=None):
        """Prefill + greedy decode (HF `generate(do_sample=False)` semantics: a sequence stops at its first EOS id; at most
        `max_new_tokens` tokens).  Returns (prefill logits [n_seq, vocab], list of n_seq python lists of generated ids, ending with the EOS id if one was produced).
        `step_logits` (a list) collects the logits of every decode step for tests."""
        import ctypes
        w = self.w
        dev = X.device
        T = X.shape[0]
        logits0 = self.prefill(X, cu_seqlens, positions, n_seq, max_len, last_rows, kv_rows=n_seq * max_new_tokens)
        m = self._c_model()
        ptrs = (ctypes.c_void_p * len(w.layers))(*[self.kv[l].data_ptr() for l in range(len(w.layers))])
        tok = torch.empty((max_new_tokens, n_seq), device=dev, dtype=torch.int32)  # tok[s] = ids produced at step s (tok[0] from the prefill)
        L.check(L.lib().d3d_argmax_rows(L.ptr(logits0), logits0.stride(0), n_seq, w.vocab, L.ptr(tok[0]), L.stream_ptr()))
        x32 = torch.empty((n_seq, w.hidden), device=dev, dtype=torch.float32)
        a16 = torch.empty((n_seq, w.hidden), device=dev, dtype=w.dtype)
        att16 = torch.empty((n_seq, w.hidden), device=dev, dtype=w.dtype)
        h16 = torch.empty((n_seq, w.ffn), device=dev, dtype=w.dtype)
        tab = torch.empty((n_seq, self.Dh), device=dev, dtype=torch.float32)
        pos = torch.empty((n_seq,), device=dev, dtype=torch.int32)
        logits = torch.empty((n_seq, w.vocab), device=dev, dtype=torch.float32)
        host = torch.empty((max_new_tokens, n_seq), dtype=torch.int32).pin_memory()
        events = []
        eos = set(int(e) for e in eos_ids)
        stream = torch.cuda.current_stream()

        def finished_upto(s):  # every sequence has produced an EOS among tok[0..s] (host copy already complete)
            h = host[: s + 1].numpy()
            return bool(eos) and all(any(int(t) in eos for t in h[:, b]) for b in range(n_seq))

        n_steps = 1
        host[0].copy_(tok[0], non_blocking=True)
        ev = torch.cuda.Event(); ev.record(stream); events.append(ev)
        for s in range(max_new_tokens - 1):
            # stop as soon as a COMPLETED earlier step shows that all sequences ended (no blocking: the GPU keeps a step or two of lead)
            done = [i for i, e in enumerate(events) if e.query()]
            if done and finished_upto(max(done)):
                break
            L.check(L.lib().d3d_lm_decode_step(ctypes.addressof(m), ctypes.cast(ptrs, ctypes.c_void_p), self.kv.stride(1), L.ptr(cu_seqlens), n_seq, T, s,
                                               L.ptr(tok[s]), L.ptr(self.inv_freq), L.ptr(x32), L.ptr(a16), L.ptr(att16), L.ptr(h16), L.ptr(tab),
                                               L.ptr(pos), L.ptr(logits), L.ptr(tok[s + 1]), L.stream_ptr()))
            if step_logits is not None:  # tests: logits that chose tok[s + 1]
                step_logits.append(logits.clone())
            host[s + 1].copy_(tok[s + 1], non_blocking=True)
            ev = torch.cuda.Event(); ev.record(stream); events.append(ev)
            n_steps += 1
        stream.synchronize()
        out = []
        hn = host[:n_steps].numpy()
        for b in range(n_seq):
            ids = []
            for t in hn[:, b].tolist():
                ids.append(int(t))  # like HF generate, the EOS id itself is part of the output
                if t in eos:
                    break
            out.append(ids)
        return logits0, out
