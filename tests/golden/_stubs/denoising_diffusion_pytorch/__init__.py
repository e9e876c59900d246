"""Stub of the third-party package `denoising-diffusion-pytorch==1.9.4` (pinned at
/root/reference/README.md:151, imported at src/core/network/nn/generator.py:3).
The real package is not installed in this image and there is no network; only the two
position-embedding classes the reference imports are restated here, following the public
lucidrains definition. Used ONLY by tests/golden/make_golden.py to import the reference."""
