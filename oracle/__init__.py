"""CPU oracle for the segmentation hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU in fp32, the arithmetic the reference performs through
TensorFlow-1.x ops on its FCN-8s hot path (`Network/model/FCN.py:52-107,117-171,334-340`).

PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures, and TensorFlow
is not installable in this image, so this oracle could not be checked against output of the
reference itself.  It is instead cross-checked against naive NumPy loop restatements
(`oracle/naive.py`) and fp64 finite differences (tests/test_oracle.py).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this package.  The product path (`semanticsegmentation_tensorflow_b200`)
never does.
"""
