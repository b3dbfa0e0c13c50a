"""CPU oracle for the FFT cross-correlation hot path -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this package.  The product
(libaudiosync_cuda.so and old-audiosync_b200/audiosync_cuda) never does.
"""
