"""Synthetic inputs of the BASELINE.json configurations, generated with the product's own ``sig_gen``:
the brain_sim 1.5 T phantom of the reference's only fixture (``data/params_brain_sim_1_5T.csv``), pseudo-noise ensemble
members, MRSI voxels with perturbed parameters (C4) and the ragged large ensemble (C5).  Large batches are synthesised on
the device (``llck_multi_fid_batched``); the seeded noise comes from torch's device generator.  Used by ``bench.py`` and by
the parity tests (which hand the very same arrays to their CPU checker)."""
import numpy as np

from . import sig_gen

DWELL = 5e-4

# data/params_brain_sim_1_5T.csv of the reference (amplitude, t2, frequency, phase), sorted by frequency like
# llckbdm/_tests/fixtures.py:23-35 does
BRAIN_SIM_PARAMS = np.array([
    [1.0, 0.002712968, 75.31704, 0.0],
    [0.11611, 0.0138504155, 160.06464, 0.0],
    [0.291727, 0.0199203187, 246.46896, 0.0],
    [0.428882, 0.0735294118, 255.5172, 0.0],
    [0.0290276, 0.0066489362, 268.8984, 0.0],
    [0.0184325, 0.0909090909, 269.5356, 0.0],
    [0.0450798, 0.0833333333, 290.43576, 0.0],
    [0.0427286, 0.1162790698, 299.99376, 0.0],
    [0.202612, 0.0925925926, 386.7804, 0.0],
    [0.0777794, 0.1136363636, 410.22936, 0.0],
    [0.0201887, 0.1052631579, 414.94464, 0.0],
    [0.0411176, 0.1470588235, 455.08824, 0.0],
    [0.0150218, 0.2222222222, 464.5188, 0.0],
    [0.105428, 0.0456621005, 482.3604, 0.0],
    [0.299129, 0.04, 503.388, 0.0],
    [0.824383, 0.0087950748, 525.30768, 0.0],
])


def brain_sim(N=2048, sigma=1e-3, seed=0, dwell=DWELL):
    """brain_sim(N, sigma, seed) of SURVEY.md §8(d): the 16-component FID on ``linspace(0, dwell N, N)`` (reference
    _tests/fixtures.py:18-20, 43-47) plus seeded complex Gaussian noise."""
    t = np.linspace(0, dwell * N, N, endpoint=False)
    c = sig_gen.multi_fid(t, BRAIN_SIM_PARAMS)
    if sigma > 0:
        rng = np.random.default_rng(seed)
        c = c + sigma * (rng.standard_normal(N) + 1j * rng.standard_normal(N))
    return c


def pseudo_noise_members(base, seeds, sigma_pn=1e-6):
    """One pseudo-noise draw of ``base`` per seed (the LLC ensemble members of README.md:10 that the reference's code never
    generates itself): list of host arrays."""
    out = []
    n = len(base)
    for s in seeds:
        g = np.random.default_rng(int(s))
        out.append(base + sigma_pn * (g.standard_normal(n) + 1j * g.standard_normal(n)))
    return out


def c2_m_range():
    """Config C2: 100 truncations m in [700, 1024] of one 2048-point FID."""
    return [700 + round(k * 324 / 99) for k in range(100)]


def c4_voxel_params(v):
    """Config C4, voxel v: amplitudes x U(0.5, 1.5), T2 x U(0.8, 1.2), F + N(0, 2 Hz), phase 0 (rng = default_rng(v))."""
    g = np.random.default_rng(int(v))
    p = BRAIN_SIM_PARAMS.copy()
    p[:, 0] *= g.uniform(0.5, 1.5, 16)
    p[:, 1] *= g.uniform(0.8, 1.2, 16)
    p[:, 2] += g.normal(0, 2.0, 16)
    return p


def c4_voxels_device(v0, v1, N=1024, sigma=1e-3, dwell=DWELL, device=None):
    """FIDs of voxels [v0, v1) of the C4 grid as a complex128 CUDA tensor [v1 - v0, N]: models synthesised on the device, noise
    from the device generator seeded with v0 (reproducible for a given (v0, v1))."""
    import torch
    sets = [c4_voxel_params(v) for v in range(v0, v1)]
    out = sig_gen.multi_fid_batched_device(sets, N, dwell, device=device)
    gen = torch.Generator(device=out.device)
    gen.manual_seed(int(v0) + 1)
    noise = torch.randn((v1 - v0, N, 2), dtype=torch.float64, device=out.device, generator=gen)
    return out + sigma * torch.view_as_complex(noise)


def c5_member_sizes(k0, k1, stride=1):
    """Config C5: m_k = 512 + (k mod 513); ``stride`` 37 spreads a short slice over the whole [512, 1024] range."""
    return [512 + (k * stride) % 513 for k in range(k0, k1)]


def c5_members_device(base_dev, k0, k1, sigma_pn=1e-6):
    """Pseudo-noise members k0..k1-1 of the C5 ensemble on the device: ``base_dev`` [N] + sigma_pn * seeded noise -> [k1-k0, N]."""
    import torch
    gen = torch.Generator(device=base_dev.device)
    gen.manual_seed(1000 + int(k0))
    noise = torch.randn((k1 - k0, base_dev.numel(), 2), dtype=torch.float64, device=base_dev.device, generator=gen)
    return base_dev.reshape(1, -1) + sigma_pn * torch.view_as_complex(noise)
