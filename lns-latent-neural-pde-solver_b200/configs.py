"""Model hyper-parameters of the four stage-2 configurations the reference ships (values restated from
configs/ns2d_stage2_prop.yml, configs/SW_stage2_prop.yml, configs/twophase_stage2_prop.yml and
configs/twophase_stage2_cond_prop.yml; only the keys the model constructors read -- paths, optimiser and logging
settings are not part of the rollout path).  ``get_config(name)`` returns an ``argparse.Namespace`` exactly like the
reference's ``dict2namespace(yaml.safe_load(...))`` would for those keys."""
import argparse
import copy

_COMMON_AE = dict(
    encoder_channels=[64, 64, 64, 128, 128], fourier_resolutions=[], encoder_res_blocks=1,
    use_fa=True, decoder_channels=[128, 128, 64, 64], decoder_res_blocks=1, final_smoothing=False,
    disable_coarse_attn=False, prop_n_embd=128,
)

CONFIGS = {
    # NS2d 64x64 vorticity, latent 16 x 8 x 8, fully periodic
    "ns2d": dict(_COMMON_AE, kind="ns2d", latent_dim=16, Ly=64, Lx=64, resolution=64, in_channels=1,
                 latent_resolution=8, is_periodic=True, use_attn_enc=False, attn_resolutions=[16, 32],
                 attn_heads=8, attn_dim=64, noise_level=0., prop_n_block=3, dilation=2,
                 val_steps=29, val_batch=32),
    # shallow water 96x192 (vx, vy, pressure), latent 64 x 12 x 24, periodic in x
    "sw": dict(_COMMON_AE, kind="sw", latent_dim=64, Ly=96, Lx=192, resolutions=[96, 192], in_channels=3,
               latent_resolution=12, periodic_direction="x", hw_ratio=2, attn_resolutions=[24, 48],
               decoder_attn_heads=8, decoder_attn_dim=64, prop_n_block=4, dilation=3,
               val_steps=42, val_batch=10),
    # two-phase tank sloshing 61x121 (vx, vy, p, vof), latent 64 x 7 x 15, zero padding
    "twophase": dict(_COMMON_AE, kind="twophase", latent_dim=64, Ly=61, Lx=121, resolutions=[61, 121], in_channels=4,
                     latent_resolution=7, is_periodic=False, hw_ratio=2, attn_resolutions=[15, 30],
                     decoder_attn_heads=8, decoder_attn_dim=64, prop_n_block=4, dilation=2,
                     val_steps=78, val_batch=32),
    # same AE, propagator conditioned on the (normalised) oscillation frequency; the YAML has no disable_coarse_attn
    # key, the reference decoder reads it anyway (modules/autoencoder2d_nonsquared.py:170) -> None
    "twophase_cond": dict(_COMMON_AE, kind="twophase_cond", latent_dim=64, Ly=61, Lx=121, resolutions=[61, 121],
                          in_channels=4, latent_resolution=7, is_periodic=False, hw_ratio=2,
                          attn_resolutions=[15, 30], decoder_attn_heads=8, decoder_attn_dim=64, cond_channels=1,
                          cond_emb_channels=64, prop_n_block=4, dilation=2, disable_coarse_attn=None,
                          val_steps=78, val_batch=32),
}


def get_config(name):
    if name not in CONFIGS:
        raise KeyError(f"unknown config {name!r}; choose from {sorted(CONFIGS)}")
    return argparse.Namespace(**copy.deepcopy(CONFIGS[name]))
