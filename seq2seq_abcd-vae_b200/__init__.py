# coding: utf-8
"""B200-native implementation of the seq2seq ABCD-VAE training / encoding hot path.

The importable name of this package contains a hyphen (it mirrors the reference repository's name), so load it
with importlib, e.g. ``abcd = importlib.import_module('seq2seq_abcd-vae_b200')``; `abcdvae_b200.py` at the
repository root does exactly that and registers the alias ``abcdvae_b200``.
"""
from . import _lib  # noqa: F401
from .modules import model, data_utils  # noqa: F401

__all__ = ['model', 'data_utils', '_lib']
