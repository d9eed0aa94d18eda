"""The simulated lidar of a fleet on the GPU: laser scans of the shared static map (second half).

`synth.raycast` (host numpy) defines the sensor model — it builds the synthetic workloads on machines without a GPU;
`raycast_gpu` runs the same model in csrc/sensor_kernel.cuh (b200mpc_raycast_batch) with the map's occupancy bits staged
in shared memory, so that a closed loop of the obstacle-active variant runs scan -> obstacle list -> solve -> control step on
the device (fleet.FleetObstacleAvoidance).  The two agree sample for sample (tests/test_gpu_parity.py)."""
import numpy as np

from .obstacles import _default_solver


def map_bits(m):
    """Occupancy bits of a map dict (synth.load_map): (H, ceil(W/32)) uint32, bit (c & 31) of word c >> 5 = cell (r, c)."""
    occ = np.asarray(m["occ"], dtype=bool)
    H, W = occ.shape
    wpr = (W + 31) // 32
    padded = np.zeros((H, wpr * 32), dtype=np.uint8)
    padded[:, :W] = occ
    return np.packbits(padded.reshape(H, wpr, 32), axis=2, bitorder="little").view(np.uint32).reshape(H, wpr)


def raycast_gpu(m, pose, n_beams=360, angle_min=0.0, angle_max=6.28, range_min=0.12, range_max=3.5, step=0.01, solver=None):
    """Scans (B,n_beams) and the (angle_min, angle_max) pair, like synth.raycast(m, pos, yaw, ...); pose (B,3)."""
    pose = np.atleast_2d(np.asarray(pose, dtype=np.float64))
    H, W = m["occ"].shape
    S = solver or _default_solver()
    scan = S.raycast_batch(map_bits(m), H, W, m["origin"], m["resolution"], pose, angle_min, angle_max, range_min, range_max,
                           step, n_beams)
    return scan, np.array([angle_min, angle_max])
