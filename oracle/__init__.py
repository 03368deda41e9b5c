"""Test-only CPU/PyTorch restatement of the reference path. See vsl_oracle.py."""
