"""b200kf -- B200-native batched GPS/IMU Kalman-filter engine.

Drop-in for the filter loop and brute-force schedule search of
IseanB/SensorFusion-KalmanFilter (kf_workers.py, hw5_2.py): the same Python call surface on
top of hand-written sm_100a CUDA kernels reached through the C ABI in include/b200kf.h.
There is no CPU fallback: every compute entry point raises if libb200kf.so or a CUDA device
is missing.
"""
__version__ = "0.1.0"
