"""puzzlenet_b200 -- B200 (sm_100a) implementation of the PuzzleNet encoder + pair-matching
forward and the approximate-EMD loss behind the reference's own Python API.

Drop-in modules (same names, signatures and error behaviour as the reference):
  puzzlenet_b200.pointnet_util   <- pointnet_util.py
  puzzlenet_b200.emd             <- PyTorchEMD/emd.py     (and puzzlenet_b200.emd_cuda <- emd_cuda)
  puzzlenet_b200.model5_b        <- model5_b.py (TouchedRegraster.predict5 / forward, encoder blocks)
  puzzlenet_b200.se3             <- se_math/se3.py (exp, transform)
`puzzlenet_b200.dropin.install()` registers them under the reference's module names.
"""
__version__ = "0.1.0"
