# coding: utf-8
"""Import shim: exposes the package directory `seq2seq_abcd-vae_b200/` under the valid module name
`abcdvae_b200` (``import abcdvae_b200 as abcd; abcd.model.ABCDSampler(...)``)."""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
if _ROOT not in sys.path:
	sys.path.insert(0, _ROOT)
_pkg = importlib.import_module('seq2seq_abcd-vae_b200')
sys.modules[__name__] = _pkg
