"""a15 input path (POL:438): `llava_processor(images=observations['rgb'])` = HF CLIPImageProcessor on the PIL path (transformers 4.46 pin):
Pillow's fixed-point two-pass bicubic resize (a = -0.5, uint8 intermediate), rescale, normalise.  The oracle restatement is pinned against
Pillow itself and against transformers' own PIL-backed processor; the CUDA kernel is then compared bit for bit with the oracle."""
import numpy as np
import pytest
import torch


def _images(seed, n, size):
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, size=(n, size, size, 3), dtype=np.uint8)
    img[0] = (np.arange(size * size * 3) % 253).reshape(size, size, 3).astype(np.uint8)  # smooth ramps
    if n > 1:
        img[1, ::2] = 255  # hard edges: overshoot of the negative lobes exercises clip8
        img[1, 1::2] = 0
    return img


@pytest.mark.parametrize("size", [224, 256, 480, 100])
def test_oracle_pil_resize_matches_pillow(size):
    from PIL import Image
    from oracle import nn_ops as NN
    img = _images(size, 3, size)
    want = np.stack([np.asarray(Image.fromarray(im).resize((336, 336), resample=Image.BICUBIC)) for im in img])
    got = NN.pil_bicubic_resize_u8(img, 336, 336)
    assert got.dtype == np.uint8 and np.array_equal(got, want)


def test_oracle_matches_transformers_pil_image_processor():
    from transformers.models.clip.image_processing_pil_clip import CLIPImageProcessorPil
    from oracle import nn_ops as NN
    proc = CLIPImageProcessorPil(size={"shortest_edge": 336}, crop_size={"height": 336, "width": 336}, do_resize=True, do_center_crop=True,
                                 do_rescale=True, do_normalize=True, do_convert_rgb=True, resample=3,
                                 image_mean=[0.48145466, 0.4578275, 0.40821073], image_std=[0.26862954, 0.26130258, 0.27577711])
    for size in (224, 336):
        img = _images(size + 1, 2, size)
        want = proc(images=torch.from_numpy(img), return_tensors="pt")["pixel_values"]
        got = NN.hf_clip_image_process(img, 336)
        assert torch.equal(got, want), size


def test_host_tables_match_the_oracle():
    from dynam3d_b200 import ops
    from oracle import nn_ops as NN
    for n_in in (224, 256, 480, 100, 336):
        b0, k0 = NN.pil_resample_tables(n_in, 336)
        b1, k1 = ops.pil_bicubic_tables(n_in, 336)
        assert np.array_equal(b0, b1) and np.array_equal(k0, k1)


@pytest.mark.gpu
@pytest.mark.parametrize("size", [224, 256, 100])
def test_cuda_pil_resize_bit_exact(size):
    from dynam3d_b200 import ops
    from oracle import nn_ops as NN
    img = _images(size + 2, 3, size)
    got = ops.pil_bicubic_resize(torch.from_numpy(img).cuda(), 336, 336).cpu().numpy()
    assert np.array_equal(got, NN.pil_bicubic_resize_u8(img, 336, 336))


@pytest.mark.gpu
def test_llava_tower_input_path_matches_hf_processor():
    """The im2col operand of the tower's patch embedding equals the HF processor's pixel_values (fp16) for 224^2 observations."""
    import torch.nn.functional as F
    from dynam3d_b200 import ops
    from oracle import nn_ops as NN
    img = _images(9, 2, 224)
    cols = ops.preprocess_im2col(ops.pil_bicubic_resize(torch.from_numpy(img).cuda(), 336, 336)).cpu().float()
    want = NN.hf_clip_image_process(img, 336).half().float()
    ref = F.unfold(want, 14, stride=14).transpose(1, 2).reshape(-1, 588)
    assert torch.equal(cols[:, :588], ref)
