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
