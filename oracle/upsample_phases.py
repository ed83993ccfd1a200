"""TEST INFRASTRUCTURE / design groundwork (not shipped on the product path).

Nearest x2 up-sampling followed by a 3x3 convolution (modules/basics.py:295-299 UpSampleBlock; the decoders'
nn.Upsample(size=2*(h, w)) -> Conv3x3, modules/autoencoder2d.py:134-136) equals four 2x2 convolutions of the SOURCE image,
one per output phase (py, px) = (Y % 2, X % 2):

    out[2y+py, 2x+px] = sum_{a,b in {0,1}}  Wp[py][px][a][b] . src[y + py - 1 + a, x + px - 1 + b]
    Wp[0][.][0] = w[0],        Wp[0][.][1] = w[1] + w[2]        (rows; the same split for columns)
    Wp[1][.][0] = w[0] + w[1], Wp[1][.][1] = w[2]

i.e. 16 filter taps per 4 output pixels instead of 36: 2.25x fewer MACs.  Padding of the up-sampled image (zeros or circular)
becomes the same padding of the source image.  The halo conv engine already addresses taps as (row, column) offsets 0..2 into a
shared-memory halo: phase (py, px) tap (a, b) is offset (py + a, px + b) of the SAME halo, only the filter slice differs --
DESIGN.md "Next", item 2.  This file pins the algebra on the CPU (tests/test_oracle.py::test_upsample_phase_decomposition)."""
import torch
import torch.nn.functional as F


def phase_filters(w):
    """w [Cout, Cin, 3, 3] -> [2, 2, Cout, Cin, 2, 2]: the 2x2 filter of every output phase (py, px)."""
    rows = [[w[:, :, 0], w[:, :, 1] + w[:, :, 2]], [w[:, :, 0] + w[:, :, 1], w[:, :, 2]]]  # [py][a] -> [Cout, Cin, 3 (kx)]
    out = w.new_zeros(2, 2, w.shape[0], w.shape[1], 2, 2)
    for py in range(2):
        for a in range(2):
            r = rows[py][a]
            cols = [[r[:, :, 0], r[:, :, 1] + r[:, :, 2]], [r[:, :, 0] + r[:, :, 1], r[:, :, 2]]]  # [px][b]
            for px in range(2):
                for b in range(2):
                    out[py, px, :, :, a, b] = cols[px][b]
    return out


def conv3x3_of_up2(x, w, bias=None, circular=(False, False)):
    """Reference: conv3x3(pad 1, zeros / circular per axis) of the nearest x2 up-sampled x."""
    up = F.interpolate(x, scale_factor=2.0, mode="nearest")
    up = F.pad(up, (1, 1, 0, 0), mode="circular" if circular[1] else "constant")
    up = F.pad(up, (0, 0, 1, 1), mode="circular" if circular[0] else "constant")
    return F.conv2d(up, w, bias)


def conv_up2_by_phases(x, w, bias=None, circular=(False, False)):
    """The same result from four 2x2 convolutions of the source image (one per output phase)."""
    B, _, H, W = x.shape
    wp = phase_filters(w)
    xp = F.pad(x, (1, 1, 0, 0), mode="circular" if circular[1] else "constant")
    xp = F.pad(xp, (0, 0, 1, 1), mode="circular" if circular[0] else "constant")  # source rows / columns -1 .. H / W
    out = x.new_zeros(B, w.shape[0], 2 * H, 2 * W)
    for py in range(2):
        for px in range(2):
            # taps (a, b) read source (y + py - 1 + a, x + px - 1 + b) = padded (y + py + a, x + px + b)
            y = F.conv2d(xp[:, :, py:py + H + 1, px:px + W + 1], wp[py, px], bias)
            out[:, :, py::2, px::2] = y
    return out
