"""``utils.dict2namespace`` -- every reference train script imports it (train_stage2_ns2d.py:16) but the reference
repository does not ship the file.  Provided so that the unmodified scripts can import from this repository."""
import argparse


def dict2namespace(config):
    namespace = argparse.Namespace()
    for key, value in config.items():
        setattr(namespace, key, dict2namespace(value) if isinstance(value, dict) else value)
    return namespace
