"""The reference imports ``modules.siren_module`` (modules/basics.py:6, modules/factorized_attention.py:8) but does
not ship it; nothing on the rollout path uses these names.  Placeholders keep ``import`` statements working."""


class SirenNet:
    def __init__(self, *a, **k):
        raise NotImplementedError("SirenNet is not part of the reference repository (missing file) and unused on the path")


class SirenWrapper(SirenNet):
    pass
