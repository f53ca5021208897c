"""ORACLE (test infrastructure only - never imported by the product path).

The ResNet-50 FE of the reference's configs is not reference code but a library model:
`torchvision.models.resnet50(pretrained=True)` with `fc = Linear(2048, 512)` (configs/dog_fe/fe_dogs_config.py:96-109).
The checker is therefore that very torchvision module in fp32 (eager PyTorch: conv2d / batch_norm / max_pool2d), built here
with seeded random weights because the pretrained checkpoint needs a download."""
import torch
import torchvision


def build(embedding=512, seed=0, layers=(3, 4, 6, 3)):
    """layers (3, 4, 6, 3) = resnet50; shallower stacks of the same Bottleneck are better conditioned test cases"""
    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    m = torchvision.models.ResNet(torchvision.models.resnet.Bottleneck, list(layers))
    m.fc = torch.nn.Linear(2048, embedding)
    with torch.no_grad():
        for name, p in m.named_parameters():
            if p.dim() == 1 and 'bn' in name or 'downsample.1' in name:
                # BatchNorm affine: away from the (1, 0) init so that gamma / beta gradients are exercised
                p.copy_(torch.rand(p.shape, generator=g) * 0.5 + 0.75 if name.endswith('weight') else torch.randn(p.shape, generator=g) * 0.1)
        for name, b in m.named_buffers():
            if name.endswith('running_mean'):
                b.copy_(torch.randn(b.shape, generator=g) * 0.1)
            elif name.endswith('running_var'):
                b.copy_(torch.rand(b.shape, generator=g) * 0.5 + 0.75)
    return m


def conv3x3_grid_reference(x_rows, weight, B, H, W):
    """3x3 / pad 1 convolution of activations given as rows of the padded grid -> rows of the padded grid (ring rows zero)"""
    Cin = x_rows.shape[1]
    x = x_rows.view(B, H + 2, W + 2, Cin)[:, 1:-1, 1:-1].permute(0, 3, 1, 2).float()
    y = torch.nn.functional.conv2d(x, weight.float(), padding=1)
    out = torch.zeros(B, H + 2, W + 2, weight.shape[0], dtype=torch.float32, device=x_rows.device)
    out[:, 1:-1, 1:-1] = y.permute(0, 2, 3, 1)
    return out.view(-1, weight.shape[0])


def _r(x):
    """round to the bf16 grid, gradient = identity (the product path stores every activation as bf16)"""
    return x.to(torch.bfloat16).float()


def forward_emulated(m, img01):
    """The training-mode forward of `m` (a torchvision ResNet of Bottlenecks) in fp32 eager PyTorch with the activations and
    the weights rounded to bf16 at exactly the points where the B200 path stores them.  A random-init ResNet in training
    mode is chaotic enough that two correct bf16 implementations disagree by tens of percent in the gradients (PyTorch's own
    autocast does, against fp32); with the rounding points matched the ReLU masks and batch statistics coincide, and what is
    left is the rounding of the gradients themselves - so this is the oracle that can check the COMPOSITION (skip connections,
    strides, down-sampling branches, the fused stem) tightly.  img01: float [B, 3, H, W] in [0, 1]."""
    F = torch.nn.functional

    def bn(mod, x):
        return F.batch_norm(x, None, None, mod.weight, mod.bias, True, 0.0, mod.eps)

    def conv(mod, x):
        return _r(F.conv2d(x, _r(mod.weight), None, mod.stride, mod.padding))

    x = _r(F.max_pool2d(torch.relu(bn(m.bn1, conv(m.conv1, _r(img01)))), 3, 2, 1))
    for layer in (m.layer1, m.layer2, m.layer3, m.layer4):
        for blk in layer:
            y1 = _r(torch.relu(bn(blk.bn1, conv(blk.conv1, x))))
            y2 = _r(torch.relu(bn(blk.bn2, conv(blk.conv2, y1))))
            a3 = conv(blk.conv3, y2)
            idn = x if blk.downsample is None else _r(bn(blk.downsample[1], conv(blk.downsample[0], x)))
            x = _r(torch.relu(bn(blk.bn3, a3) + idn))
    pooled = _r(x.mean((2, 3)))
    return pooled @ _r(m.fc.weight).t() + m.fc.bias
