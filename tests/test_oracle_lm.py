"""Pins oracle/nn_ops.lm_prefill against HF transformers' LlamaForCausalLM (the third-party implementation the
reference calls through LlavaForConditionalGeneration, POL:123-127) on a small config, CPU fp32."""
import pytest
import torch


def test_lm_prefill_matches_hf_llama():
    transformers = pytest.importorskip("transformers")
    from dynam3d_b200 import synth
    from oracle import nn_ops as NN
    hidden, layers, heads, ffn, vocab = 192, 2, 2, 384, 500  # head_dim 96 like Phi-3-mini
    cfg = transformers.LlamaConfig(hidden_size=hidden, intermediate_size=ffn, num_hidden_layers=layers, num_attention_heads=heads,
                                   num_key_value_heads=heads, vocab_size=vocab, rms_norm_eps=1e-5, rope_theta=10000.0,
                                   max_position_embeddings=4096, attention_bias=False, mlp_bias=False, tie_word_embeddings=False)
    model = transformers.LlamaForCausalLM(cfg).eval()
    sd = synth.lm_state_dict(3, hidden, layers, ffn, vocab)
    missing = model.load_state_dict(sd, strict=False)
    assert not [k for k in missing.missing_keys if "rotary" not in k and "inv_freq" not in k], missing
    lens = [37, 5]
    emb = synth.hash_uniform((sum(lens), hidden), 11, 1.0)
    want = []
    s = 0
    with torch.no_grad():
        for n in lens:
            out = model(inputs_embeds=emb[s:s + n][None])
            want.append(out.logits[0, -1])
            s += n
    got = NN.lm_prefill(emb, lens, sd, layers, heads)
    assert (torch.stack(want) - got).abs().max().item() < 1e-4


def test_greedy_decode_matches_hf_generate():
    """Pins the cache-free decode restatement (oracle/nn_ops.lm_greedy_decode / lm_teacher_forced_logits) against HF
    `generate(inputs_embeds=..., do_sample=False)` -- the call the reference makes at POL:463 -- with and without an EOS id."""
    transformers = pytest.importorskip("transformers")
    from dynam3d_b200 import synth
    from oracle import nn_ops as NN
    hidden, layers, heads, ffn, vocab = 192, 2, 2, 384, 500
    cfg = transformers.LlamaConfig(hidden_size=hidden, intermediate_size=ffn, num_hidden_layers=layers, num_attention_heads=heads,
                                   num_key_value_heads=heads, vocab_size=vocab, rms_norm_eps=1e-5, rope_theta=10000.0,
                                   max_position_embeddings=4096, attention_bias=False, mlp_bias=False, tie_word_embeddings=False)
    model = transformers.LlamaForCausalLM(cfg).eval()
    sd = synth.lm_state_dict(3, hidden, layers, ffn, vocab)
    # larger logit margins than the 0.02-std init gives, so fp32 summation-order noise cannot flip an arg-max
    sd["lm_head.weight"] = sd["lm_head.weight"] * 20.0
    model.load_state_dict(sd, strict=False)
    lens = [23, 9]
    emb = synth.hash_uniform((sum(lens), hidden), 12, 1.0)
    n_new = 6
    hf, s = [], 0
    with torch.no_grad():
        for n in lens:
            out = model.generate(inputs_embeds=emb[s:s + n][None], attention_mask=torch.ones(1, n, dtype=torch.long), max_new_tokens=n_new,
                                 do_sample=False, pad_token_id=0, eos_token_id=None)
            hf.append(out[0].tolist())
            s += n
    got = NN.lm_greedy_decode(emb, lens, sd, layers, heads, max_new_tokens=n_new)
    assert got == hf, (got, hf)
    # teacher-forced logits of the same tokens: every step's arg-max is the token HF chose
    lg = NN.lm_teacher_forced_logits(emb, lens, [t[:n_new - 1] for t in hf], sd, layers, heads)
    assert lg.shape[0] == n_new and lg.argmax(-1).t().tolist() == hf
    # EOS: HF keeps the EOS id as the last token of the sequence that produced it
    eos = hf[0][2]
    with torch.no_grad():
        out = model.generate(inputs_embeds=emb[:lens[0]][None], attention_mask=torch.ones(1, lens[0], dtype=torch.long), max_new_tokens=n_new,
                             do_sample=False, pad_token_id=0, eos_token_id=eos)
    cut = NN.lm_greedy_decode(emb[:lens[0]], [lens[0]], sd, layers, heads, max_new_tokens=n_new, eos_ids=(eos,))
    assert cut[0] == out[0].tolist() and cut[0][-1] == eos
