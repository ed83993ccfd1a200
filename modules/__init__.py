"""Drop-in replacements for the reference's ``modules`` package (same file names, class names, constructor
arguments and ``state_dict`` layouts).  Parameters live in ordinary ``torch.nn`` containers so
reference checkpoints load with ``strict=True``; ``forward`` runs the sm_100a kernels of liblns_b200.so through
``lns_b200.ops`` (inference only, CUDA only, no PyTorch-eager fallback)."""
