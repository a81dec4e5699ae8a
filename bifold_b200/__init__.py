"""bifold_b200 — B200-native (sm_100a) implementation of the BiFold policy forward/backward hot path.

Drop-in factories mirroring the reference package layout:
    bifold_b200.models.Models / Components   <- bifold.models
    bifold_b200.losses.Losses                <- bifold.losses
    bifold_b200.optim.Optimizers/Schedulers  <- bifold.optim
The computation runs in libbifold_b200.so (bifold_b200/csrc, C ABI in include/bifold_b200.h).
"""
__version__ = "0.1.0"
