"""Input pipeline: the oracle against the reference's own torchvision transforms on PIL images, and the host
parameter sampler against the reference's RNG draw order."""
import pytest
import torch


def _images(seed=0):
    g = torch.Generator().manual_seed(seed)
    out = []
    for h, w in ((300, 400), (224, 224), (500, 333), (97, 180)):
        # smooth content + noise: resampling differences show on both
        yy, xx = torch.meshgrid(torch.linspace(0, 6.28, h), torch.linspace(0, 6.28, w), indexing="ij")
        base = torch.stack([torch.sin(yy) * torch.cos(xx), torch.sin(2 * xx), torch.cos(yy + xx)], dim=-1) * 100 + 128
        out.append((base + torch.randn(h, w, 3, generator=g) * 20).clamp(0, 255).to(torch.uint8))
    return out


def _reference_pipeline(size, scale):
    """datasets/augmentations.py:36-73 with the RandAugment stage removed."""
    from torchvision import transforms as T
    return T.Compose([
        T.RandomResizedCrop(size, scale=scale, interpolation=T.InterpolationMode.BICUBIC),
        T.RandomHorizontalFlip(p=0.5),
        T.ToTensor(),
        T.Normalize(mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225)),
        T.RandomErasing(p=0.25, scale=(0.02, 1 / 3), ratio=(0.3, 3.3), inplace=False),
    ])


@pytest.mark.parametrize("seed", [0, 1, 2, 3, 4, 5])
def test_sampler_and_oracle_match_the_reference_transforms_on_pil(seed):
    from PIL import Image
    from dinov2_distillation_b200.input_pipeline import sample_params
    from oracle import augment_ref
    size, scale = 64, (0.32, 1.0)
    imgs = _images(seed)
    tf = _reference_pipeline(size, scale)
    torch.manual_seed(100 + seed)
    want = torch.stack([tf(Image.fromarray(im.numpy())) for im in imgs])
    torch.manual_seed(100 + seed)
    crop, flip, erase = sample_params([(im.shape[0], im.shape[1]) for im in imgs], size, scale)
    got = augment_ref.augment_batch(imgs, crop, flip, erase, size)
    # same draws (crop box, flip, erase box) and Pillow's integer resize restated exactly: equal bit for bit
    assert torch.equal(got, want), (got - want).abs().max().item()


def test_sampler_shapes_and_erase_fallback():
    from dinov2_distillation_b200.input_pipeline import sample_params
    torch.manual_seed(0)
    crop, flip, erase = sample_params([(50, 60)] * 200, 32, (0.08, 1.0))
    assert crop.shape == (200, 4) and flip.shape == (200,) and erase.shape == (200, 4)
    assert ((crop[:, 0] + crop[:, 2] <= 50) & (crop[:, 1] + crop[:, 3] <= 60) & (crop[:, 2] > 0) & (crop[:, 3] > 0)).all()
    assert 60 < int(flip.sum()) < 140
    n_er = int((erase[:, 2] > 0).sum())
    assert 20 < n_er < 85                                   # p = 0.25
    e = erase[erase[:, 2] > 0]
    assert ((e[:, 0] + e[:, 2] <= 32) & (e[:, 1] + e[:, 3] <= 32)).all()
