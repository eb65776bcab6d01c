"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference modules.

Loads `Feature_Fields` (Dynam3D_VLN/vlnce_baselines/models/feature_fields.py)
and the OpenAI-CLIP `VisionTransformer`
(Dynam3D_VLN/vlnce_baselines/models/encoders/clip/model.py) straight from the
read-only reference tree, in THIS container only, so that
`oracle/make_golden.py` can (a) validate the restatement in `oracle/` and
(b) write golden vectors under `tests/golden/`.

Nothing in the product package (`dynam3d_b200/`), `bench.py` or the `-m gpu`
tests imports this file: `/root/reference` does not exist on the GPU box.

The reference needs four third-party modules that are not installed here; they
are replaced by minimal stand-ins injected into `sys.modules`:

* `torch_kdtree.build_kd_tree` (feature_fields.py:7,246,606) -> exact
  brute-force K-NN, squared L2 in fp32, ascending, lowest index on ties.
* `open3d` (feature_fields.py:8) -> the three entry points of the posed-dataset branch (Image,
  PinholeCameraIntrinsic, PointCloud.create_from_depth_image) restated from open3d 0.19's published
  algorithm; the habitat branch never touches them.
* `configargparse` (feature_fields.py:24) -> argparse.
* `vlnce_baselines.models.fastsam` (feature_fields.py:17) -> stub classes; the
  segmentation (FastSAM output) is an INPUT of the hot path, so
  `get_patch_segm` is replaced by a function returning the synthetic map.

* `torch.cuda.get_device_properties` / `torch.cuda.memory_allocated` (free-memory probe in the merge branch,
  feature_fields.py:678-680) are patched to report 80 GB / 0 B when no CUDA device exists, which selects the
  same (grad-enabled) encoder call the reference makes on a GPU with > 10 GB free.

Two NumPy-1.x-only comparisons (`ndarray == []`, feature_fields.py:557,567) are
rewritten in memory to `len(...) == 0`; NumPy >= 1.25 raises on them.
"""
import argparse
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF_ROOT = os.environ.get("DYNAM3D_REFERENCE", "/root/reference")
FF_PATH = os.path.join(REF_ROOT, "Dynam3D_VLN/vlnce_baselines/models/feature_fields.py")
CLIP_MODEL_PATH = os.path.join(REF_ROOT, "Dynam3D_VLN/vlnce_baselines/models/encoders/clip/model.py")
PRETRAIN_CLIP_MODEL_PATH = os.path.join(REF_ROOT, "Dynam3D_Pretrain/src_3dff/models/encoders/clip/model.py")


def reference_available():
    return os.path.isfile(FF_PATH) and os.path.isfile(CLIP_MODEL_PATH)


class _BruteForceTree:
    """Stand-in for torch_kdtree's tree object (query semantics used at feature_fields.py:606,610)."""

    def __init__(self, points):
        self.points = points.detach().to(torch.float32).clone()

    def query(self, q, nr_nns_searches=1):
        k = int(nr_nns_searches)
        q = q.detach().to(torch.float32)
        n_q = q.shape[0]
        if k == 0:
            return (torch.zeros((n_q, 0), dtype=torch.float32), torch.zeros((n_q, 0), dtype=torch.int64))
        p = self.points.numpy()
        qq = q.numpy()
        dx = (qq[:, None, 0] - p[None, :, 0]).astype(np.float32)
        dy = (qq[:, None, 1] - p[None, :, 1]).astype(np.float32)
        dz = (qq[:, None, 2] - p[None, :, 2]).astype(np.float32)
        d2 = ((dx * dx).astype(np.float32) + (dy * dy).astype(np.float32)).astype(np.float32)
        d2 = (d2 + (dz * dz).astype(np.float32)).astype(np.float32)
        order = np.argsort(d2, axis=1, kind="stable")[:, :k]
        dist = np.take_along_axis(d2, order, axis=1)
        return torch.from_numpy(dist), torch.from_numpy(order.astype(np.int64))


def _open3d_stand_in():
    """The three open3d entry points the posed-dataset branch uses (feature_fields.py:50-60, 262-273), following open3d 0.19's
    published algorithm (ImageFactory.cpp ConvertDepthToFloatImage, PointCloudFactory.cpp CreatePointCloudFromFloatDepthImage):
    float z = depth / depth_scale, z >= depth_trunc -> 0, pixels with z <= 0 are DROPPED, x = (u - cx) z / fx in double."""
    o3d = types.ModuleType("open3d")
    geometry, camera = types.ModuleType("open3d.geometry"), types.ModuleType("open3d.camera")

    class Image:
        def __init__(self, array):
            self.array = np.asarray(array)

    class PinholeCameraIntrinsic:
        def __init__(self, width, height, fx, fy, cx, cy):
            self.width, self.height, self.fx, self.fy, self.cx, self.cy = int(width), int(height), float(fx), float(fy), float(cx), float(cy)

    class PointCloud:
        def __init__(self):
            self.points = np.zeros((0, 3), np.float64)

        def __iadd__(self, other):
            self.points = np.concatenate([self.points, other.points], 0)
            return self

        @staticmethod
        def create_from_depth_image(depth, intrinsic, depth_scale=1000.0, depth_trunc=1000.0, stride=1):
            z = depth.array.astype(np.float32)  # uint16 -> float (CreateFloatImage) or already float
            z = (z / np.float32(depth_scale)).astype(np.float32)
            z = np.where(z >= np.float32(depth_trunc), np.float32(0), z)
            H, W = z.shape
            v, u = np.nonzero(z > 0)  # row-major scan order, invalid pixels dropped
            zz = z[v, u].astype(np.float64)
            pc = PointCloud()
            pc.points = np.stack([(u - intrinsic.cx) * zz / intrinsic.fx, (v - intrinsic.cy) * zz / intrinsic.fy, zz], -1)
            return pc

    geometry.Image, geometry.PointCloud, camera.PinholeCameraIntrinsic = Image, PointCloud, PinholeCameraIntrinsic
    o3d.geometry, o3d.camera = geometry, camera
    return o3d


def _install_stubs():
    if "torch_kdtree" not in sys.modules:
        m = types.ModuleType("torch_kdtree")
        m.build_kd_tree = lambda pts: _BruteForceTree(pts)
        sys.modules["torch_kdtree"] = m
    if "open3d" not in sys.modules:
        sys.modules["open3d"] = _open3d_stand_in()
    if "configargparse" not in sys.modules:
        m = types.ModuleType("configargparse")
        m.ArgumentParser = argparse.ArgumentParser
        sys.modules["configargparse"] = m
    for name in ("vlnce_baselines", "vlnce_baselines.models"):
        if name not in sys.modules:
            pkg = types.ModuleType(name)
            pkg.__path__ = []
            sys.modules[name] = pkg
    if "vlnce_baselines.models.fastsam" not in sys.modules:
        m = types.ModuleType("vlnce_baselines.models.fastsam")

        class FastSAM:  # noqa: D401 - stub
            def __init__(self, *a, **k):
                pass

        class FastSAMPrompt:
            def __init__(self, *a, **k):
                pass

        m.FastSAM = FastSAM
        m.FastSAMPrompt = FastSAMPrompt
        sys.modules["vlnce_baselines.models.fastsam"] = m


def _patch_cuda_probe():
    if torch.cuda.is_available() or getattr(torch.cuda, "_d3d_probe_patched", False):
        return
    props = types.SimpleNamespace(total_memory=80 * 1024 ** 3)
    torch.cuda.get_device_properties = lambda *a, **k: props
    torch.cuda.memory_allocated = lambda *a, **k: 0
    torch.cuda._d3d_probe_patched = True


_FF_MODULE = None


def load_reference_feature_fields_module():
    """Exec the reference feature_fields.py (with the two NumPy-2 rewrites) and return the module."""
    global _FF_MODULE
    if _FF_MODULE is not None:
        return _FF_MODULE
    _install_stubs()
    _patch_cuda_probe()
    with open(FF_PATH, "r") as f:
        src = f.read()
    n1 = src.count("if self.global_patch_position[b] == []:")
    n2 = src.count("if self.global_patch_fts[b] == []:")
    assert n1 == 1 and n2 == 1, "reference source changed; update the in-memory rewrite"
    src = src.replace("if self.global_patch_position[b] == []:", "if len(self.global_patch_position[b]) == 0:")
    src = src.replace("if self.global_patch_fts[b] == []:", "if len(self.global_patch_fts[b]) == 0:")
    mod = types.ModuleType("ref_feature_fields")
    mod.__file__ = FF_PATH
    exec(compile(src, FF_PATH, "exec"), mod.__dict__)
    _FF_MODULE = mod
    return mod


def make_reference_feature_fields(batch_size=1, seed=0):
    """Instantiate the reference Feature_Fields on CPU with seeded default init."""
    mod = load_reference_feature_fields_module()
    argv = sys.argv
    sys.argv = [argv[0]]
    try:
        torch.manual_seed(seed)
        ff = mod.Feature_Fields(batch_size=batch_size, device="cpu")
    finally:
        sys.argv = argv
    ff.eval()
    return ff


def attach_segmentation(ff, segm_fn):
    """Replace FastSAM (`get_patch_segm`, feature_fields.py:400-430) by `segm_fn(batch_image)->int64 [N,24,24]`."""
    ff.get_patch_segm = lambda batch_image, *a, **k: segm_fn(batch_image)


def load_reference_clip_model_module(pretrain=False):
    path = PRETRAIN_CLIP_MODEL_PATH if pretrain else CLIP_MODEL_PATH
    spec = importlib.util.spec_from_file_location("ref_clip_model_pre" if pretrain else "ref_clip_model", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


# ------------------------------------------------------------------------------------------------
# Pretrain renderer (a18): Dynam3D_Pretrain/src_3dff/models/feature_fields.py with a tinycudann stand-in
# ------------------------------------------------------------------------------------------------
PFF_PATH = os.path.join(REF_ROOT, "Dynam3D_Pretrain/src_3dff/models/feature_fields.py")


class _TcnnNetwork(torch.nn.Module):
    """Stand-in for `tinycudann.Network` with otype CutlassMLP (PFF:221-243; tinycudann 2.0 is CUDA-only and absent): bias-free fully
    connected layers, fp16 weights / activations with fp32 accumulation, LeakyReLU slope 0.01, output width padded to a multiple of 16,
    ONE flat fp32 `params` vector holding the row-major [out, in] matrices in layer order (the published layout).  The arithmetic of
    tinycudann itself therefore stays unpinned; everything AROUND it in render_view_3d_patch is the reference's own code."""

    def __init__(self, n_input_dims, n_output_dims, network_config, seed=1337):
        super().__init__()
        assert network_config["otype"] == "CutlassMLP"
        self.n_in, self.n_out = int(n_input_dims), int(n_output_dims)
        self.width, self.n_hidden = int(network_config["n_neurons"]), int(network_config["n_hidden_layers"])
        self.act, self.out_act = network_config["activation"], network_config["output_activation"]
        pad = (self.n_out + 15) // 16 * 16
        self.dims = [(self.width, self.n_in)] + [(self.width, self.width)] * (self.n_hidden - 1) + [(pad, self.width)]
        n = sum(o * i for o, i in self.dims)
        g = torch.Generator().manual_seed(seed)
        self.params = torch.nn.Parameter((torch.rand(n, generator=g) * 2 - 1) * (3.0 / self.width) ** 0.5)

    def forward(self, x):
        h = x.to(torch.float16).to(torch.float32)
        off = 0
        for li, (o, i) in enumerate(self.dims):
            w = self.params[off: off + o * i].view(o, i).to(torch.float16).to(torch.float32)
            off += o * i
            h = h @ w.t()
            last = li == len(self.dims) - 1
            act = self.out_act if last else self.act
            if act == "LeakyReLU":
                h = torch.nn.functional.leaky_relu(h, 0.01)
            else:
                assert act == "None", act
            h = h.to(torch.float16).to(torch.float32)
        return h[:, : self.n_out].to(torch.float16)


_PFF_MODULE = None


def load_reference_pretrain_ff_module():
    """Exec the reference Pretrain feature_fields.py unmodified, with stand-ins for tinycudann / torch_kdtree / open3d / configargparse / FastSAM."""
    global _PFF_MODULE
    if _PFF_MODULE is not None:
        return _PFF_MODULE
    _install_stubs()
    _patch_cuda_probe()
    if "tinycudann" not in sys.modules:
        m = types.ModuleType("tinycudann")
        m.Network = _TcnnNetwork
        sys.modules["tinycudann"] = m
    for name in ("src_3dff", "src_3dff.models"):
        if name not in sys.modules:
            pkg = types.ModuleType(name)
            pkg.__path__ = []
            sys.modules[name] = pkg
    if "src_3dff.models.fastsam" not in sys.modules:
        sys.modules["src_3dff.models.fastsam"] = sys.modules["vlnce_baselines.models.fastsam"]
    with open(PFF_PATH, "r") as f:
        src = f.read()
    mod = types.ModuleType("ref_pretrain_feature_fields")
    mod.__file__ = PFF_PATH
    exec(compile(src, PFF_PATH, "exec"), mod.__dict__)
    _PFF_MODULE = mod
    return mod


def make_reference_pretrain_feature_fields(batch_size=1, seed=0):
    mod = load_reference_pretrain_ff_module()
    argv = sys.argv
    sys.argv = [argv[0]]
    try:
        torch.manual_seed(seed)
        ff = mod.Feature_Fields(batch_size=batch_size, device="cpu")
    finally:
        sys.argv = argv
    ff.eval()
    return ff


def reference_render_view(ff, patch_pos, patch_dir, patch_scale, patch_fts16, position_hab, heading):
    """Run the reference's `render_view_3d_patch` (PFF:494-625, habitat mode) on a given patch cloud of episode 0.  The reference executes
    this under fp16 autocast on the GPU (`.to(torch.float16)` inputs into fp32 nn.Linear weights, PFF:479-483); on CPU the same autocast
    region is opened for float16."""
    ff.batch_size, ff.mode = 1, "habitat"
    ff.gt_pcd_tree = None
    ff.global_patch_fts = [np.asarray(patch_fts16, np.float16)]
    ff.global_patch_directions = [np.asarray(patch_dir, np.float32)]
    ff.global_patch_scales = [np.asarray(patch_scale, np.float32)]
    ff.global_patch_position = [torch.from_numpy(np.asarray(patch_pos, np.float32))]
    ff.patch_tree = [ff.get_patch_tree(0)]
    ff.sampled_rays = ff.get_rays_habitat()
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.float16):
        fts, pos, _ = ff.render_view_3d_patch(batch_position=[np.asarray(position_hab, np.float32).copy()], batch_heading=[float(heading)])
    return fts[0].reshape(-1, fts.shape[-1]).float().numpy(), pos[0].reshape(-1, 3).float().numpy()


def load_reference_waypoint_predictor():
    """The reference's candidate-waypoint predictor, unmodified: returns (TRM_net module, waypoint_pred.utils module).  Its vendored
    `pytorch_transformer` package wants boto3 (download helpers, unused) and the pip name `pytorch_transformers` (TRM_net.py:7): both are
    satisfied with empty stand-ins / an alias of the vendored modeling_bert."""
    _install_stubs()
    import importlib.machinery

    def absent(top):
        return top not in sys.modules and importlib.util.find_spec(top) is None

    def stand_in(name, package=False):
        m = types.ModuleType(name)
        m.__spec__ = importlib.machinery.ModuleSpec(name, None, is_package=package)  # with a spec: importlib.util.find_spec() on it must not raise
        if package:
            m.__path__ = []
        sys.modules[name] = m
        return m

    if absent("boto3"):
        stand_in("boto3")
    if absent("botocore"):
        stand_in("botocore", package=True)
        stand_in("botocore.exceptions").ClientError = Exception
    base = os.path.join(REF_ROOT, "Dynam3D_VLN", "vlnce_baselines", "waypoint_pred")
    pk = "vlnce_baselines.waypoint_pred"
    for name, sub in ((pk, ""), (pk + ".transformer", "transformer"), (pk + ".transformer.pytorch_transformer", os.path.join("transformer", "pytorch_transformer"))):
        if name not in sys.modules:
            pkg = types.ModuleType(name)
            pkg.__path__ = [os.path.join(base, sub)]
            sys.modules[name] = pkg

    def load(name, rel):
        if name in sys.modules and getattr(sys.modules[name], "__file__", None):
            return sys.modules[name]
        spec = importlib.util.spec_from_file_location(name, os.path.join(base, rel))
        m = importlib.util.module_from_spec(spec)
        sys.modules[name] = m
        spec.loader.exec_module(m)
        return m

    pt = pk + ".transformer.pytorch_transformer."
    load(pt + "file_utils", os.path.join("transformer", "pytorch_transformer", "file_utils.py"))
    load(pt + "modeling_utils", os.path.join("transformer", "pytorch_transformer", "modeling_utils.py"))
    mb = load(pt + "modeling_bert", os.path.join("transformer", "pytorch_transformer", "modeling_bert.py"))
    if "pytorch_transformers" not in sys.modules:
        alias = types.ModuleType("pytorch_transformers")
        alias.__spec__ = importlib.machinery.ModuleSpec("pytorch_transformers", None)
        alias.BertConfig = mb.BertConfig
        sys.modules["pytorch_transformers"] = alias
    utils = load(pk + ".utils", "utils.py")
    sys.modules[pk].utils = utils
    load(pk + ".transformer.waypoint_bert", os.path.join("transformer", "waypoint_bert.py"))
    trm = load(pk + ".TRM_net", "TRM_net.py")
    return trm, utils
