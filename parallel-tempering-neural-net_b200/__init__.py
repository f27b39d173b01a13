"""B200-native parallel-tempering sampler for Bayesian feed-forward networks.

The product is ``csrc/libptfnn.so`` (hand-written sm_100a CUDA behind the C ABI of
``include/ptfnn.h``); the modules here are the thin Python host side:

  capi            ctypes binding of the C ABI (fails loudly if the library is missing)
  sampler         ``Sampler``: one handle = the temperatures held by one GPU
  regression      ``Network`` / ``ptReplica`` / ``ParallelTempering`` with the signatures of
                  multicore-pt-regression/pt_timeseries_regression.py
  classification  the same for multicore-pt-classification/pt_classification.py
  distributed     ladder partitioned over ranks (torch.distributed) with boundary swaps
"""
__version__ = "0.1.0"
