"""Minimal stand-in for the `xfuser` package surface scripts/inference/generate.py:218-229 and
wan/text2video.py:91 touch.  The reference delegates Ulysses/ring attention to xfuser (un-vendored third-party
code); here sequence parallelism is implemented in wan/distributed/ulysses.py, and this shim only keeps the
process-group bookkeeping API."""
