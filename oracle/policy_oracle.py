"""TEST INFRASTRUCTURE ONLY -- CPU restatement of `Dynam3D_VLN.forward` (eval branch) up to the next-action logits.

Follows Dynam3D_VLN/vlnce_baselines/models/Policy_Dynam3D_VLN.py (= POL): depth handling POL:336-341,350; CLIP encode
POL:343-345 (ENC:267-284, CLIPM:219-238); 3D memory POL:349-363 (via ff_oracle); projections POL:83-111,432-435; LLaVA
image features POL:441-453 (HF CLIPVisionModel hidden_states[-2] + multi_modal_projector); splice POL:456; LM prefill of
`llava.generate` POL:463 (HF Llama/Phi-3 math, pinned in tests/test_oracle_lm.py).

Parity status: the 3D-memory part is PINNED against the reference run here; `Policy_Dynam3D_VLN.py` itself cannot be imported
(needs habitat / gym / peft), so this file is a restatement of its forward anchored on the pinned sub-blocks: "parity unpinned"
for the glue (prompt splice, projector wiring), pinned for every numerical sub-block.
Q13: for num_of_views > 1 the LLM sees view 0's patch tokens.
The LLaVA tower's input follows POL:438: HF CLIPImageProcessor (Pillow bicubic, nn_ops.hf_clip_image_process -- pinned against Pillow and
transformers' own PIL-backed processor in tests/test_image_processor.py), NOT the torchvision transform of the CLIPEncoder path.
"""
import numpy as np
import torch

from . import geometry as G
from . import nn_ops as NN
from .ff_oracle import FeatureFieldsOracle


def hf_clip_to_openai(sd, prefix):
    g = lambda k: sd[prefix + k]
    o = {"conv1.weight": g("embeddings.patch_embedding.weight"), "class_embedding": g("embeddings.class_embedding"),
         "positional_embedding": g("embeddings.position_embedding.weight"),
         "ln_pre.weight": g("pre_layrnorm.weight"), "ln_pre.bias": g("pre_layrnorm.bias")}
    n_layers = len({k[len(prefix):].split(".")[2] for k in sd if k.startswith(prefix + "encoder.layers.")})
    for l in range(n_layers):
        s, d = f"encoder.layers.{l}.", f"transformer.resblocks.{l}."
        o[d + "ln_1.weight"], o[d + "ln_1.bias"] = g(s + "layer_norm1.weight"), g(s + "layer_norm1.bias")
        o[d + "ln_2.weight"], o[d + "ln_2.bias"] = g(s + "layer_norm2.weight"), g(s + "layer_norm2.bias")
        o[d + "attn.in_proj_weight"] = torch.cat([g(s + f"self_attn.{x}_proj.weight") for x in "qkv"], 0)
        o[d + "attn.in_proj_bias"] = torch.cat([g(s + f"self_attn.{x}_proj.bias") for x in "qkv"], 0)
        o[d + "attn.out_proj.weight"], o[d + "attn.out_proj.bias"] = g(s + "self_attn.out_proj.weight"), g(s + "self_attn.out_proj.bias")
        o[d + "mlp.c_fc.weight"], o[d + "mlp.c_fc.bias"] = g(s + "mlp.fc1.weight"), g(s + "mlp.fc1.bias")
        o[d + "mlp.c_proj.weight"], o[d + "mlp.c_proj.bias"] = g(s + "mlp.fc2.weight"), g(s + "mlp.fc2.bias")
    return o, n_layers


class PolicyOracle:
    def __init__(self, policy_sd, clip_sd, llava_sd, clip_layers=24, clip_heads=16, lm_layers=32, lm_heads=32, batch_size=1,
                 rnd=None, lm_rnd=None, q1_fix=False, q7_fix=False):
        self.P = {k: v.detach().float() for k, v in policy_sd.items()}
        ff_sd = {k[len("feature_fields."):]: v for k, v in self.P.items() if k.startswith("feature_fields.")}
        self.ff = FeatureFieldsOracle(ff_sd, batch_size=batch_size, rnd=rnd, q7_fix=q7_fix)
        self.clip_sd, self.clip_layers, self.clip_heads = clip_sd, clip_layers, clip_heads
        self.llava = llava_sd
        self.tower_sd, self.tower_layers = hf_clip_to_openai(llava_sd, "vision_tower.vision_model.")
        self.lm_layers, self.lm_heads = lm_layers, lm_heads
        self.rnd, self.lm_rnd = rnd, lm_rnd if lm_rnd is not None else rnd
        self.q1_fix = q1_fix
        self.history = [["none\n"] * 4 for _ in range(batch_size)]

    def _tokens(self, fts, rel, pos_prefix, proj_prefix):
        if len(fts) == 0:
            return torch.zeros((0, 3072))
        pe = NN.mlp_ln_gelu(torch.from_numpy(rel), self.P, pos_prefix, self.rnd)
        return NN.mlp_ln_gelu(torch.cat([torch.from_numpy(fts), pe], -1), self.P, proj_prefix, self.rnd)

    def step_logits(self, obs, agent_positions, agent_headings, input_ids, num_of_views=1, delete_old_features=True):
        """obs: dict rgb u8 [B*V,H,W,3], depth f32 [B*V,Hd,Wd,1], patch_segm [B,V,24,24]; input_ids: per-episode prompt ids, or a
        callable (b, n_image_tokens) -> ids (the prompt needs the 3D-token counts, POL:436, which only exist after the memory update)."""
        B, V = self.ff.batch_size, num_of_views
        depth = np.asarray(obs["depth"], np.float32)
        d576 = G.depth_patch_grid(depth, B, V, q1_fix=self.q1_fix)
        x = NN.clip_preprocess(obs["rgb"], 336, rnd=self.rnd)
        _, grid = NN.vit_forward(x, self.clip_sd, self.clip_layers, self.clip_heads, rnd=self.rnd)
        grid = grid.numpy().astype(np.float16).reshape(B, V, 576, 768)  # POL:345 / FF:500
        if delete_old_features:
            full = G.preprocess_depth(depth, (0.0, 10.0)).reshape(B, V, depth.shape[1], depth.shape[2])
            self.ff.delete_old_features_from_camera_frustum(full, agent_positions, agent_headings, num_of_views=V)
        self.ff.update_feature_fields(d576, grid, np.asarray(obs["patch_segm"]), agent_positions, agent_headings, num_of_views=V)
        env = self.ff.get_environment_features(agent_positions, agent_headings)
        info = G.patch_3d_info(d576.reshape(B * V, -1))
        embeds, lens = [], []
        emb_table = self.llava["language_model.model.embed_tokens.weight"].float()
        for b in range(B):
            v0 = b * V
            feat6 = torch.from_numpy(np.stack([info[0][v0], info[1][v0], info[2][v0], np.sin(info[3][v0]), np.cos(info[3][v0]), info[4][v0]], -1))
            patch_pos = NN.mlp_ln_gelu(feat6, self.P, "patch_position_embedding", self.rnd)
            # POL:438: the tower's pixel_values come from the HF processor (Pillow resize), cast to fp16 whatever the model precision
            px = NN.hf_clip_image_process(np.asarray(obs["rgb"])[v0:v0 + 1], 336, rnd=NN.round_fp16)
            hid = NN.vit_forward(px, self.tower_sd, self.tower_layers, self.clip_heads, rnd=self.rnd,
                                 n_layers_run=self.tower_layers - 1, return_hidden=True)[0, 1:]
            h = NN.gelu(NN.linear(hid, self.llava["multi_modal_projector.linear_1.weight"], self.llava["multi_modal_projector.linear_1.bias"], self.rnd))
            patch = NN.linear(h, self.llava["multi_modal_projector.linear_2.weight"], self.llava["multi_modal_projector.linear_2.bias"], self.rnd)
            patch = patch + patch_pos
            inst = self._tokens(env["batch_instance_fts"][b], env["batch_instance_relative_position"][b], "instance_position_embedding", "instance_projector")
            zone = self._tokens(env["batch_zone_fts"][b], env["batch_zone_relative_position"][b], "zone_position_embedding", "zone_projector")
            n_img = 576 + len(inst) + len(zone)
            ids = torch.tensor(list(input_ids(b, n_img) if callable(input_ids) else input_ids[b]), dtype=torch.long)
            e = emb_table[ids]
            seq = torch.cat([e[:2], patch, inst, zone, e[n_img + 2:]], 0)  # POL:456
            embeds.append(seq)
            lens.append(len(seq))
        lm_sd = {k[len("language_model."):]: v for k, v in self.llava.items() if k.startswith("language_model.")}
        self.last_lens = lens
        return NN.lm_prefill(torch.cat(embeds, 0), lens, lm_sd, self.lm_layers, self.lm_heads, rnd=self.lm_rnd)
