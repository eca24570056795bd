"""B200 (sm_100a) backend: ctypes binding of libflashb200.so + autograd ops + denoiser engines."""
