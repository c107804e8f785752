"""Layer-level entry points of the drop-in surface on the GPU -- ConvBlock.forward (down / up / none, with
FiLM vector, conditioning input, residual and target length), PReLU_Conv.forward with the PReLU
activation, film() -- against the oracle's restatement of blocks.py on the same weights.  The networks do
not call these (they run the fused program); LoRA-style consumers and users of the module API do."""
import pytest
import torch

from common import rel_rms

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 4e-3      # 16-bit storage of inputs / outputs of every fused stage (kernel tests: 3e-3 per conv)


def sd_of(mod, prefix="m"):
    return {f"{prefix}.{k}": v.detach().clone() for k, v in mod.state_dict().items()}


def randomize(mod, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in mod.named_parameters():
            if name.endswith("prelu.weight"):
                p.copy_(0.1 + 0.3 * torch.rand(p.shape, generator=g))
            elif "bias" in name.rsplit(".", 1)[-1]:
                p.copy_(0.1 * torch.randn(p.shape, generator=g))
    return mod


@pytest.mark.parametrize("direction,rate,antialias,C,T", [
    ("none", None, False, 64, 333),
    ("down", 2, True, 32, 1001),      # odd length: right-pad to the stride multiple (blocks.py:205-210)
    ("down", 4, False, 64, 800),
    ("up", 4, True, 64, 250),
    ("up", 5, False, 128, 60),
    ("none", None, False, 256, 77),   # per-conv path (no fused trunk at this width)
])
@pytest.mark.parametrize("with_cond", [False, True])
def test_conv_block_forward_vs_oracle(direction, rate, antialias, C, T, with_cond):
    from open_universe_b200.networks.universe.blocks import ConvBlock
    from oracle.universe_oracle import conv_block
    torch.manual_seed(5)
    blk = randomize(ConvBlock(C, rate_change=rate, rate_change_dir=direction, antialiasing=antialias,
                              use_weight_norm=True), 6)
    g = torch.Generator().manual_seed(7)
    B = 2
    cin = 2 * C if direction == "up" else C
    h = torch.randn(B, cin, T, generator=g)
    length = res = None
    if direction == "up":
        length = T * rate - 1                     # skip tensor one sample shorter than the upsampled length
        res = torch.randn(B, C, length, generator=g)
    elif direction == "none" and with_cond:
        res = torch.randn(B, C, T, generator=g)   # rate-preserving block with a residual (score.py:196-210)
    t_mid = length if length is not None else T
    noise_cond = torch.randn(B, 2 * C, generator=g) if with_cond else None
    input_cond = torch.randn(B, C, t_mid, generator=g) if with_cond else None
    with torch.no_grad():
        want = conv_block(sd_of(blk), "m", h, direction, rate, antialias, noise_cond, input_cond, res, length)
    blk = blk.to(DEV)
    dev = lambda t: None if t is None else t.to(DEV)   # noqa: E731
    got = blk(dev(h), dev(noise_cond), dev(input_cond), dev(res), length)
    for name, a, b in zip(("h_out", "skip", "cond_out"), got, want):
        assert a.shape == b.shape, (name, a.shape, b.shape)
        err = rel_rms(a.cpu(), b)
        assert err < TOL, (name, err)


@pytest.mark.parametrize("cin,cout,k,stride,transpose,antialias,T", [
    (32, 32, 5, 1, False, False, 500),
    (64, 128, 4, 4, False, True, 403),
    (128, 64, 4, 4, True, True, 100),
    (48, 96, 2, 2, False, False, 301),
])
def test_prelu_conv_forward_vs_oracle(cin, cout, k, stride, transpose, antialias, T):
    from open_universe_b200.networks.universe.blocks import PReLU_Conv
    from oracle.universe_oracle import prelu_conv
    torch.manual_seed(11)
    pc = randomize(PReLU_Conv(cin, cout, k, stride=stride, padding="same" if stride == 1 else 0,
                              use_transpose=transpose, use_weight_norm=True, use_antialiasing=antialias), 12)
    x = torch.randn(2, cin, T, generator=torch.Generator().manual_seed(13))
    with torch.no_grad():
        want = prelu_conv(sd_of(pc), "m", x, stride=stride, transpose=transpose, same=stride == 1,
                          antialias=antialias)
    got = pc.to(DEV)(x.to(DEV)).cpu()
    assert got.shape == want.shape
    assert rel_rms(got, want) < TOL


def test_film_kernel_vs_oracle():
    from open_universe_b200.networks.universe.blocks import film
    from oracle.universe_oracle import film as film_ref
    g = torch.Generator().manual_seed(3)
    x, y = torch.randn(3, 48, 1001, generator=g), torch.randn(3, 96, generator=g)
    got = film(x.to(DEV), y.to(DEV)).cpu()
    assert torch.allclose(got, film_ref(x, y), rtol=1e-6, atol=1e-6)
    with pytest.raises(ValueError):
        film(x.to(DEV), y[:, :50].to(DEV))
