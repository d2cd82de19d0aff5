"""GPU input pipeline (csrc/augment.cu) against the oracle (which tests/test_oracle_augment.py pins bit for bit
on the reference's torchvision / PIL transforms)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _images(seed, sizes):
    g = torch.Generator().manual_seed(seed)
    out = []
    for h, w in sizes:
        yy, xx = torch.meshgrid(torch.linspace(0, 6.28, h), torch.linspace(0, 6.28, w), indexing="ij")
        base = torch.stack([torch.sin(yy) * torch.cos(xx), torch.sin(2 * xx), torch.cos(yy + xx)], dim=-1) * 100 + 128
        out.append((base + torch.randn(h, w, 3, generator=g) * 20).clamp(0, 255).to(torch.uint8))
    return out


@pytest.mark.parametrize("seed,size", [(0, 64), (1, 224), (2, 37)])
def test_augment_batch_bit_exact_with_oracle(seed, size):
    """Down- and up-scaling crops, odd image sizes, flips, erase boxes: equal to the oracle bit for bit (integer resize,
    IEEE division for ToTensor / Normalize)."""
    from dinov2_distillation_b200.input_pipeline import augment_batch, sample_params
    from oracle import augment_ref
    imgs = _images(seed, ((300, 400), (224, 224), (500, 333), (97, 180), (31, 45), (640, 480), (224, 224), (1000, 700)))
    torch.manual_seed(7 + seed)
    crop, flip, erase = sample_params([(im.shape[0], im.shape[1]) for im in imgs], size, (0.08, 1.0), erase_p=0.6)
    assert int(flip.sum()) not in (0, len(imgs)) and int((erase[:, 2] > 0).sum()) > 0
    got = augment_batch(imgs, crop, flip, erase, size).cpu()
    want = augment_ref.augment_batch(imgs, crop, flip, erase, size)
    assert got.shape == (len(imgs), 3, size, size)
    assert torch.equal(got, want), (got - want).abs().max().item()


def test_augment_identity_crop_is_a_plain_normalise():
    """Crop = whole image at the output size: the resize degenerates to a copy (weights 1, 0, 0, ...)."""
    from dinov2_distillation_b200.input_pipeline import augment_batch
    img = _images(3, ((48, 48),))[0]
    z = torch.zeros(1, 4, dtype=torch.int32)
    out = augment_batch([img], torch.tensor([[0, 0, 48, 48]], dtype=torch.int32), torch.zeros(1, dtype=torch.int32), z, 48).cpu()
    want = img.permute(2, 0, 1).float().div(255).sub(torch.tensor(augment_mean()).view(3, 1, 1)).div(torch.tensor(augment_std()).view(3, 1, 1))
    assert torch.equal(out[0], want)


def augment_mean():
    from dinov2_distillation_b200.input_pipeline import IMAGENET_DEFAULT_MEAN
    return IMAGENET_DEFAULT_MEAN


def augment_std():
    from dinov2_distillation_b200.input_pipeline import IMAGENET_DEFAULT_STD
    return IMAGENET_DEFAULT_STD


def test_gpu_augment_object_and_errors():
    from dinov2_distillation_b200._lib import B200Error
    from dinov2_distillation_b200.input_pipeline import GpuAugment, augment_batch
    imgs = _images(4, ((120, 90), (64, 200), (224, 224)))
    torch.manual_seed(0)
    out = GpuAugment((0.32, 1.0), 112)(imgs)
    assert out.shape == (3, 3, 112, 112) and out.is_cuda and torch.isfinite(out).all()
    z = torch.zeros(1, 4, dtype=torch.int32)
    with pytest.raises(ValueError):
        augment_batch(imgs[:1], torch.tensor([[0, 0, 121, 90]], dtype=torch.int32), torch.zeros(1, dtype=torch.int32), z, 32)
    with pytest.raises(B200Error):
        augment_batch(imgs[:1], torch.tensor([[0, 0, 120, 90]], dtype=torch.int32), torch.zeros(1, dtype=torch.int32), z, 32,
                      device="cpu")
