"""Checkpoint contract of the drop-in policy (ss_trainer_Dynam3D.py:75-84 saves `policy.state_dict()`, TR:214,219 restore it with
`policy.load_state_dict(ckpt["state_dict"], strict=False)`): `net.llava.*` (the only weights VLN training changes, POL:152-157) and
`net.rgb_encoder.model.*` must be nn.Module state under the reference's key names -- loaded by the unchanged trainer call, saved again."""
import numpy as np
import pytest
import torch


def _trainer_checkpoint(seed, clip_layers, lm_layers, ddp=False, small=True):
    """A synthetic `ckpt["state_dict"]` with the key layout the reference trainer writes."""
    from dynam3d_b200 import synth
    kw = dict(clip_width=128, lm_hidden=96, lm_ffn=192, vocab=320) if small else {}
    pre = "net.module." if ddp else "net."
    sd = {}
    for k, v in synth.policy_state_dict(seed).items():
        sd[pre + k] = v
    for k, v in synth.llava_state_dict(seed, clip_layers=clip_layers, lm_layers=lm_layers, lm_round_to=torch.bfloat16, **kw).items():
        sd[pre + "llava." + k] = v.to(torch.bfloat16) if v.dim() >= 2 else v  # POL:125 loads the model in bfloat16
    for k, v in synth.vit_state_dict(seed, width=128 if small else 1024, layers=clip_layers).items():
        sd[pre + "rgb_encoder.model.visual." + k] = v.to(torch.float16) if v.dim() >= 2 else v  # CLIPM:389-410 converts to fp16
    # the rest of the OpenAI CLIP model that ENC:262 keeps (text tower), and the waypoint branch's depth encoder
    sd[pre + "rgb_encoder.model.token_embedding.weight"] = torch.randn(7, 8)
    sd[pre + "rgb_encoder.model.transformer.resblocks.0.ln_1.weight"] = torch.randn(8)
    sd[pre + "rgb_encoder.model.logit_scale"] = torch.tensor(4.6)
    sd[pre + "depth_encoder.visual_encoder.backbone.conv1.0.weight"] = torch.randn(4, 1, 3, 3)
    return sd


def test_trainer_checkpoint_round_trip_cpu():
    from dynam3d_b200.policy import Policy_Dynam3D_VLN
    ckpt = _trainer_checkpoint(3, clip_layers=2, lm_layers=2)
    policy = Policy_Dynam3D_VLN.from_config(device="cpu")
    res = policy.load_state_dict(ckpt, strict=False)  # TR:219, unchanged
    assert all(k.startswith("net.depth_encoder.") for k in res.unexpected_keys), res.unexpected_keys
    assert res.missing_keys == []
    saved = policy.state_dict()  # TR:78
    for k, v in ckpt.items():
        if k.startswith("net.depth_encoder."):
            continue
        assert k in saved, k
        assert saved[k].dtype == v.dtype and torch.equal(saved[k].cpu(), v), k
    # a second load copies in place (same Parameter objects, new values) and moves the store's version
    p = policy.net.llava.get("language_model.lm_head.weight")
    v0 = policy.net.llava.version
    ckpt2 = {k: (v + 1 if k.endswith("lm_head.weight") else v) for k, v in ckpt.items()}
    policy.load_state_dict(ckpt2, strict=False)
    assert policy.net.llava.get("language_model.lm_head.weight") is p and policy.net.llava.version > v0
    assert torch.equal(p.detach(), ckpt2["net.llava.language_model.lm_head.weight"])
    # shape mismatches are errors, like for any module
    bad = dict(ckpt)
    bad["net.llava.language_model.lm_head.weight"] = torch.zeros(3, 3)
    with pytest.raises(RuntimeError, match="size mismatch"):
        policy.load_state_dict(bad, strict=False)


def test_ddp_checkpoint_and_explicit_loader_prefixes_cpu():
    from dynam3d_b200.policy import Dynam3D_VLN
    ckpt = _trainer_checkpoint(4, clip_layers=1, lm_layers=1, ddp=True)
    net = Dynam3D_VLN(device="cpu")
    net.load_policy_state_dict({k: v for k, v in ckpt.items() if "depth_encoder" not in k})  # `net.module.` prefix stripped
    assert len(net.llava) > 0 and net.rgb_encoder.model.get("visual.conv1.weight") is not None
    w = ckpt["net.module.feature_fields.instance_merge_discriminator.0.weight"]
    assert torch.equal(net.feature_fields.instance_merge_discriminator[0].weight.detach(), w)
    with pytest.raises(KeyError):
        net.load_policy_state_dict({"something.else": torch.zeros(1)})
    with pytest.raises(RuntimeError, match="missing keys"):  # strict is honoured for the projection MLPs as well
        net.load_policy_state_dict({k: v for k, v in ckpt.items() if "zone_projector.3.bias" not in k and "depth_encoder" not in k}, strict=True)


@pytest.mark.gpu
def test_trainer_load_gives_the_same_logits_as_the_explicit_loaders():
    from dynam3d_b200 import synth
    from dynam3d_b200.policy import Dynam3D_VLN, Policy_Dynam3D_VLN
    seed, V = 5, 1
    pol_sd = synth.policy_state_dict(seed)
    clip_sd = synth.vit_state_dict(seed, layers=2)
    llava_sd = synth.llava_state_dict(seed, clip_layers=2, lm_layers=2, lm_round_to=torch.float16)
    ckpt = {"net." + k: v for k, v in pol_sd.items()}
    ckpt.update({"net.llava." + k: v for k, v in llava_sd.items()})
    ckpt.update({"net.rgb_encoder.model.visual." + k: v for k, v in clip_sd.items()})
    ckpt["net.depth_encoder.x.weight"] = torch.zeros(2)
    policy = Policy_Dynam3D_VLN.from_config(q1_fix=True)
    res = policy.load_state_dict(ckpt, strict=False)
    assert all(k.startswith("net.depth_encoder.") for k in res.unexpected_keys) and res.missing_keys == []
    ref = Dynam3D_VLN(q1_fix=True)
    ref.load_policy_state_dict(pol_sd)
    ref.rgb_encoder.load_openai_state_dict(clip_sd)
    ref.llava.load_state_dict(llava_sd, max_images=1)
    ep = synth.make_episode(seed, n_steps=1, num_views=V, rgb_size=224, n_seg=16, seg_kind="voronoi")[0]
    obs = {"rgb": torch.from_numpy(ep["rgb"]), "depth": torch.from_numpy(ep["depth"]), "patch_segm": ep["segm"][None]}
    outs = []
    for net in (policy.net, ref):
        net.tokenize = synth.ToyTokenizer()
        net.feature_fields.reset(1)
        outs.append(net.forward_logits(obs, [synth.make_instruction(seed)], [ep["position"]], [ep["heading"]], num_of_views=V).cpu())
    assert torch.equal(outs[0], outs[1])
    # saving and re-loading through a fresh policy keeps them (TR:75-84 -> TR:219)
    again = Policy_Dynam3D_VLN.from_config(q1_fix=True)
    again.load_state_dict(policy.state_dict(), strict=False)
    again.net.tokenize = synth.ToyTokenizer()
    again.net.feature_fields.reset(1)
    lg = again.net.forward_logits(obs, [synth.make_instruction(seed)], [ep["position"]], [ep["heading"]], num_of_views=V).cpu()
    assert torch.equal(lg, outs[0])
