"""Argument sets of the golden op cases in tests/golden/ops.npz (the same dicts oracle/pin_reference.py used to record them)."""
FIR_KW = {
    'fir_conv0': dict(padding=[1, 1, 1, 1], gain=4),
    'fir_up2': dict(up=2, padding=[2, 1, 2, 1], gain=4),
    'fir_down2': dict(down=2, padding=[1, 1, 1, 1]),
    'fir_down2_bwd_of_up2': dict(down=2, padding=[1, 2, 1, 2], flip_filter=True, gain=4),
    'fir_crop_flip': dict(padding=[-1, 2, 0, -2], flip_filter=True, gain=0.5),
    'fir_updown_xy': dict(up=[2, 3], down=[3, 2], padding=[3, 2, 4, 1]),
    'fir_sep8': dict(up=2, padding=[4, 3, 4, 3], gain=4),
    'fir_identity': dict(up=2, padding=1),
}
CONV_KW = {
    'conv_up2': dict(up=2, padding=1, flip_weight=False),
    'conv_up2_grouped': dict(up=2, padding=1, groups=2, flip_weight=False),
    'conv_plain': dict(padding=1),
    'conv_1x1': dict(),
    'conv_1x1_up2': dict(up=2),
}
