"""Drop-in shim: with `<repo>/dropin:<repo>` in front of the reference on PYTHONPATH, the reference's own import
`from tokenizer.tokenizer_image.vq_model import ...` (sample_t2i.py:15) resolves here and re-exports the controlar_b200 implementation.
(No __init__.py on purpose: `autoregressive`, `tokenizer`, `utils` stay namespace packages, so every module this repo does
not replace — autoregressive/sample/*, dataset/*, language/* ... — still resolves to the reference tree.)"""
import controlar_b200.tokenizer.tokenizer_image.vq_model as _impl
from controlar_b200.tokenizer.tokenizer_image.vq_model import *  # noqa: F401,F403

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith('__')})   # private helpers too (e.g. find_multiple)
