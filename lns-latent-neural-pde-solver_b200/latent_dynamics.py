"""``LatentDynamics`` -- the model-assembly class the reference defines inside each stage-2 script
(train_stage2_ns2d.py:90-158, train_stage2_SW.py:91-159, train_stage2_twophase.py:91-159,
train_stage2_twophase_conditional.py:124-193): frozen autoencoder + latent propagator, ``predict`` = encode ->
K autoregressive propagator steps -> decode.  Same attribute names (``vq_ae`` / ``ae``, ``propagator``), hence the
same state_dict keys; ``predict`` runs through ``lns_b200.rollout.Rollout``."""
import torch
import torch.nn as nn


class LatentDynamics(nn.Module):
    _MAX_ROLLOUTS = 4

    def __init__(self, args, kind=None):
        super().__init__()
        kind = kind or getattr(args, "kind", None)
        if kind not in ("ns2d", "sw", "twophase", "twophase_cond"):
            raise ValueError("kind must be one of ns2d, sw, twophase, twophase_cond")
        self.kind = kind
        from modules.propagator import SimpleCNN, CondSimpleCNN
        if kind == "ns2d":
            from modules.autoencoder2d import SimpleAutoencoder
            self.vq_ae = SimpleAutoencoder(args)
        elif kind == "sw":
            from modules.autoencoder2d_half_periodic import SimpleAutoencoder
            self.vq_ae = SimpleAutoencoder(args)
        elif kind == "twophase":
            from modules.autoencoder2d_nonsquared import SimpleAutoencoder
            self.vq_ae = SimpleAutoencoder(args)
        else:
            from modules.autoencoder2d_nonsquared import SimpleAutoencoder
            self.ae = SimpleAutoencoder(args)
        self.latent_resolution = args.latent_resolution
        self.latent_dim = args.latent_dim
        if kind == "twophase_cond":
            self.propagator = CondSimpleCNN(latent_dim=args.latent_dim, cond_emb_dim=args.latent_dim,
                                            prop_n_block=args.prop_n_block, prop_n_embd=args.prop_n_embd,
                                            dilation=args.dilation)
        else:
            self.propagator = SimpleCNN(
                latent_dim=args.latent_dim, prop_n_block=args.prop_n_block, prop_n_embd=args.prop_n_embd,
                dilation=args.dilation,
                padding_mode="circular" if kind == "ns2d" else "zeros",
                periodic_direction="x" if kind == "sw" else None)
        self._rollouts = {}  # (batch, steps, to_x, device, precision) -> Rollout, at most _MAX_ROLLOUTS, LRU

    def __getstate__(self):
        """copy.deepcopy / pickle / torch.save(model): the cached rollout engines (CUDA graphs, streams, static buffers) stay
        behind; the copy builds its own on first use"""
        state = self.__dict__.copy()
        state["_rollouts"] = {}
        return state

    @property
    def autoencoder(self):
        return self.ae if self.kind == "twophase_cond" else self.vq_ae

    def load_autoencoder(self, args):
        ae = self.autoencoder
        ae.load_checkpoint(args.pretrained_checkpoint_path)
        for p in ae.parameters():
            p.requires_grad = False
        ae.eval()

    @torch.no_grad()
    def x_to_z(self, x):
        return self.autoencoder.encode(x)

    @torch.no_grad()
    def z_to_x(self, z):
        return self.autoencoder.decode(z)

    def forward(self, z_in, z_out, *args):
        """Reference signature: forward(z_in, z_out, loss_fn) (train_stage2_ns2d.py:126-141): teacher-free rollout training --
        t_out autoregressive propagator steps from z_in [b, 1, c, h, w], loss_fn(z_pred [b, t_out, c, h, w], z_out).  The rollout
        and its back-propagation run on the library's kernels as one autograd node (lns_b200.train)."""
        loss_fn = args[-1]
        param = args[0] if self.kind == "twophase_cond" else None   # forward(z_in, z_out, param, loss_fn) (conditional script :160)
        if z_in.shape[1] != 1:
            raise ValueError("z_in must hold exactly one input frame: [b, 1, c, h, w]")
        from .train import rollout_train
        z_pred = rollout_train(self.propagator, z_in[:, 0], z_out.shape[1], param=param)
        return loss_fn(z_pred, z_out)

    @torch.no_grad()
    def predict(self, x, steps, *args, to_x=False, **kw):
        """Reference signature: predict(x, steps, to_x=False) / predict(x, steps, param, to_x=False).
        Returns [B, steps, C, Ly, Lx] (to_x) or [B, steps, Cz, h, w], fp32."""
        from .rollout import Rollout
        param = None
        if self.kind == "twophase_cond":
            param = args[0] if args else kw["param"]
        elif args:
            to_x = args[0]
        if x.dim() == 5:  # the SW / conditional scripts squeeze a singleton time axis (train_stage2_SW.py:144)
            x = x.squeeze(1)
        from . import ops
        key = (x.shape[0], steps, bool(to_x), x.device, ops.get_precision())
        ro = self._rollouts.pop(key, None)
        if ro is None:
            ro = Rollout(self, batch=x.shape[0], steps=steps, to_x=to_x, use_graph=False)
            while len(self._rollouts) >= self._MAX_ROLLOUTS:  # bounded: least recently used engine (and its buffers) goes first
                self._rollouts.pop(next(iter(self._rollouts)))
        self._rollouts[key] = ro  # (re-)insert as most recently used
        out = ro(x, param) if param is not None else ro(x)
        # the engine returns its static buffer; the reference returns a fresh tensor per call (torch.stack), so does this API
        return out.clone()
