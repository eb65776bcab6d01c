"""Checkpoint sub-trees as real `nn.Module` state under the checkpoint's own key names.

The reference trainer saves `policy.state_dict()` (ss_trainer_Dynam3D.py:75-84) and restores it with
`policy.load_state_dict(ckpt["state_dict"], strict=False)` (TR:214,219).  The weights VLN training changes live under
`net.llava.*` (POL:152-157 freezes everything else); `net.rgb_encoder.model.*` is the OpenAI CLIP model (ENC:262).  The engine
runs those networks from fused, 16-bit engine-layout copies, so the modules that own them keep the checkpoint tensors in a
`WeightStore`: a tree of parameter containers that

  * adopts ANY key below its prefix when a state dict is loaded through the parent module (no architecture description or
    pre-allocated 8 GB of placeholders needed -- a freshly constructed policy accepts the trainer's checkpoint as it is);
  * copies in place when the key already exists (shape-checked), so repeated loads / `requires_grad` flags / `.to()` behave like
    any other module;
  * returns everything it holds from `state_dict()` under the same names, so TR:75-84 round-trips the weights;
  * bumps `version` on every change; the owner rebuilds its engine-layout weights lazily when the version moved.
"""
import torch
import torch.nn as nn


class _Node(nn.Module):
    """Interior container: the root store loads the whole sub-tree, the nodes themselves do nothing on load."""

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        return


class WeightStore(nn.Module):
    def __init__(self, device=None):
        super().__init__()
        self._store_device = device
        self.version = 0

    # -- explicit API ------------------------------------------------------------------------------
    def _target_device(self):
        if self._store_device is not None:
            return torch.device(self._store_device)
        for p in self.parameters():
            return p.device
        return torch.device("cuda" if torch.cuda.is_available() else "cpu")

    def _find(self, name, create=False):
        mod = self
        parts = name.split(".")
        for p in parts[:-1]:
            nxt = mod._modules.get(p)
            if nxt is None:
                if not create:
                    return None, parts[-1]
                nxt = _Node()
                mod.add_module(p, nxt)
            mod = nxt
        return mod, parts[-1]

    def put(self, name, tensor):
        """Adopt `tensor` under the dotted `name` (moved to the store's device; no copy when it already lives there)."""
        mod, leaf = self._find(name, create=True)
        t = tensor.detach().to(self._target_device())
        mod._parameters[leaf] = nn.Parameter(t, requires_grad=False)
        self.version += 1

    def get(self, name):
        mod, leaf = self._find(name)
        return None if mod is None else mod._parameters.get(leaf)

    def adopt(self, sd, prefix=""):
        """Adopt every tensor of `sd` under `prefix + key`."""
        for k, v in sd.items():
            self.put(prefix + k, v)

    def tensors(self):
        """{dotted name: tensor} of everything held (the state dict without the owner's prefix)."""
        return {k: p.detach() for k, p in self.named_parameters()}

    def __len__(self):
        return sum(1 for _ in self.parameters())

    # -- nn.Module load protocol: the root handles its whole sub-tree -----------------------------------
    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        seen = set()
        for k, v in state_dict.items():
            if not k.startswith(prefix) or not torch.is_tensor(v):
                continue
            name = k[len(prefix):]
            seen.add(name)
            cur = self.get(name)
            if cur is None:
                self.put(name, v)
            elif tuple(cur.shape) != tuple(v.shape):
                error_msgs.append(f"size mismatch for {k}: copying a param with shape {tuple(v.shape)} from checkpoint, the shape in current model is "
                                  f"{tuple(cur.shape)}.")
            else:
                with torch.no_grad():
                    cur.copy_(v)
                self.version += 1
        if strict:
            for name, _ in self.named_parameters():
                if name not in seen:
                    missing_keys.append(prefix + name)
