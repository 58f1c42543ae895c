"""Oracle: the identity loss of find_direction.py:179-180  (TEST INFRASTRUCTURE ONLY).

Restates ``id_loss/id_loss.py:8-39`` (IDLoss: pool to 256 -> crop [35:223, 32:220] -> pool to 112 -> IR-SE50 -> 1 - <f(y_hat), f(y)>, mean over
the batch, f(y) detached), ``id_loss/model_irse.py:10-49`` (Backbone, input 112, 50 layers, mode 'ir_se') and ``id_loss/helpers.py:17-119``
(get_blocks, SEModule, bottleneck_IR_SE, l2_norm) as a functional network over a flat parameter dict whose keys are the reference's
``state_dict`` names -- ``pin_reference.py`` loads this dict into the REAL ``Backbone`` (strict) and compares outputs and gradients.
The ArcFace checkpoint (``id_loss/model_ir_se50.pth``) is not in the tree: parameters are random, BatchNorm statistics included.
"""
import torch
import torch.nn.functional as F


def irse50_units():
    """(in_channel, depth, stride) of the 24 bottleneck units, helpers.py:28-38 (num_layers 50)."""
    units = []
    for cin, depth, n in ((64, 64, 3), (64, 128, 4), (128, 256, 14), (256, 512, 3)):
        units += [(cin, depth, 2)] + [(depth, depth, 1)] * (n - 1)
    return units


def _bn_params(g, p, key, c):
    p[key + '.weight'] = 1 + 0.1 * torch.randn(c, generator=g)
    p[key + '.bias'] = 0.1 * torch.randn(c, generator=g)
    p[key + '.running_mean'] = 0.1 * torch.randn(c, generator=g)
    p[key + '.running_var'] = 1 + 0.2 * torch.rand(c, generator=g)
    p[key + '.num_batches_tracked'] = torch.tensor(0)


def random_irse50_params(seed=0):
    """Random parameters under the reference's state_dict keys (He-style conv scales so that activations stay O(1) through 50 layers)."""
    g = torch.Generator().manual_seed(seed)
    p = {}
    conv = lambda o, i, k: torch.randn(o, i, k, k, generator=g) * (2.0 / (i * k * k)) ** 0.5
    p['input_layer.0.weight'] = conv(64, 3, 3)
    _bn_params(g, p, 'input_layer.1', 64)
    p['input_layer.2.weight'] = 0.25 + 0.05 * torch.randn(64, generator=g)
    for u, (cin, depth, stride) in enumerate(irse50_units()):
        b = f'body.{u}.'
        if cin != depth:
            p[b + 'shortcut_layer.0.weight'] = conv(depth, cin, 1)
            _bn_params(g, p, b + 'shortcut_layer.1', depth)
        _bn_params(g, p, b + 'res_layer.0', cin)
        p[b + 'res_layer.1.weight'] = conv(depth, cin, 3)
        p[b + 'res_layer.2.weight'] = 0.25 + 0.05 * torch.randn(depth, generator=g)
        p[b + 'res_layer.3.weight'] = conv(depth, depth, 3) * 0.5
        _bn_params(g, p, b + 'res_layer.4', depth)
        p[b + 'res_layer.5.fc1.weight'] = conv(depth // 16, depth, 1)
        p[b + 'res_layer.5.fc2.weight'] = conv(depth, depth // 16, 1)
    _bn_params(g, p, 'output_layer.0', 512)
    p['output_layer.3.weight'] = torch.randn(512, 512 * 7 * 7, generator=g) * (1.0 / (512 * 7 * 7)) ** 0.5
    p['output_layer.3.bias'] = 0.02 * torch.randn(512, generator=g)
    _bn_params(g, p, 'output_layer.4', 512)
    return p


def _bn(x, p, key):
    """BatchNorm in eval mode (id_loss.py:14 ``facenet.eval()``): the running statistics."""
    return F.batch_norm(x, p[key + '.running_mean'].to(x.dtype), p[key + '.running_var'].to(x.dtype), p[key + '.weight'].to(x.dtype),
                        p[key + '.bias'].to(x.dtype), False, 0.0, 1e-5)


def backbone_features(p, x):
    """Backbone.forward (model_irse.py:45-49) WITHOUT the final l2_norm: [N, 3, 112, 112] -> [N, 512]."""
    w = lambda k: p[k].to(x.dtype)
    x = F.prelu(_bn(F.conv2d(x, w('input_layer.0.weight'), padding=1), p, 'input_layer.1'), w('input_layer.2.weight'))
    for u, (cin, depth, stride) in enumerate(irse50_units()):
        b = f'body.{u}.'
        if cin == depth:
            shortcut = x[:, :, ::stride, ::stride]                                    # MaxPool2d(1, stride), helpers.py:98
        else:
            shortcut = _bn(F.conv2d(x, w(b + 'shortcut_layer.0.weight'), stride=stride), p, b + 'shortcut_layer.1')
        r = _bn(x, p, b + 'res_layer.0')
        r = F.prelu(F.conv2d(r, w(b + 'res_layer.1.weight'), padding=1), w(b + 'res_layer.2.weight'))
        r = _bn(F.conv2d(r, w(b + 'res_layer.3.weight'), stride=stride, padding=1), p, b + 'res_layer.4')
        s = r.mean(dim=(2, 3), keepdim=True)                                          # SEModule, helpers.py:58-76
        s = torch.sigmoid(F.conv2d(F.relu(F.conv2d(s, w(b + 'res_layer.5.fc1.weight'))), w(b + 'res_layer.5.fc2.weight')))
        x = r * s + shortcut
    x = _bn(x, p, 'output_layer.0').flatten(1)                                        # Dropout is the identity in eval mode
    x = F.linear(x, w('output_layer.3.weight'), w('output_layer.3.bias'))
    return F.batch_norm(x, p['output_layer.4.running_mean'].to(x.dtype), p['output_layer.4.running_var'].to(x.dtype),
                        w('output_layer.4.weight'), w('output_layer.4.bias'), False, 0.0, 1e-5)


def face_crop(x):
    """id_loss.py:18-22: AdaptiveAvgPool2d(256) unless the input is 256 px, crop [35:223, 32:220], AdaptiveAvgPool2d(112)."""
    if x.shape[2] != 256:
        x = F.adaptive_avg_pool2d(x, (256, 256))
    return F.adaptive_avg_pool2d(x[:, :, 35:223, 32:220], (112, 112))


def extract_feats(p, x):
    f = backbone_features(p, face_crop(x))
    return f / f.norm(dim=1, keepdim=True)                                            # l2_norm, helpers.py:17-20


def id_loss(p, y_hat, y):
    """IDLoss.forward (id_loss.py:26-39): mean over the batch of 1 - <f(y_hat), f(y)>, f(y) detached."""
    fy = extract_feats(p, y).detach()
    fh = extract_feats(p, y_hat)
    return (1.0 - (fh * fy).sum(dim=1)).mean()
