"""Regenerates tests/golden/{barc_ss_laps,tracks}.npz from the reference's DATA fixtures.

Run in the build container only (needs /root/reference):  python tests/golden/make_fixtures.py
Sources (data, not code):
  src/mpc/racing_mpc/test_data/barc_ss/ss_lap_{1,2,3}_{x,u,k,t}.txt   recorded BARC LMPC laps
  src/vehicle_dynamics_models/racing_trajectory/test_data/barc/02_barc_center.txt, 15_barc_optm.txt
  src/vehicle_dynamics_models/racing_trajectory/test_data/putnam/10_putnam_optm.txt, mgkt_optm.txt
Track tables keep only what the synthetic-input generator needs: abscissa (col 6), speed (4),
centre-line curvature (periodic cubic spline of cols 0-1), and the signed lateral offsets of the left/right boundary points
(racing_trajectory.cpp:64-79: left = +|p - p_left|, right = -|p - p_right|), total length (col 7 row 0).
"""
import os
import numpy as np

REF = "/root/reference/src"
OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    laps = {}
    d = f"{REF}/mpc/racing_mpc/test_data/barc_ss"
    for i in (1, 2, 3):
        for k in "xukt":
            laps[f"lap{i}_{k}"] = np.loadtxt(f"{d}/ss_lap_{i}_{k}.txt", ndmin=2)
    np.savez_compressed(f"{OUT}/barc_ss_laps.npz", **laps)

    tracks = {}
    td = f"{REF}/vehicle_dynamics_models/racing_trajectory/test_data"
    for name, path in (("barc_center", "barc/02_barc_center.txt"), ("barc_optm", "barc/15_barc_optm.txt"),
                       ("putnam_optm", "putnam/10_putnam_optm.txt"), ("mgkt_optm", "mgkt_optm.txt")):
        t = np.loadtxt(f"{td}/{path}")
        p = t[:, 0:2]
        tracks[f"{name}_s"] = t[:, 6]
        tracks[f"{name}_speed"] = t[:, 4]
        # curvature of the centre line from a periodic cubic spline through (s -> x, y); the table's
        # own column 5 is not a curvature in these files.
        from scipy.interpolate import CubicSpline
        L = t[0, 7] + t[0, 6]
        se = np.concatenate([t[:, 6], [L]])
        cx = CubicSpline(se, np.concatenate([p[:, 0], p[:1, 0]]), bc_type="periodic")
        cy = CubicSpline(se, np.concatenate([p[:, 1], p[:1, 1]]), bc_type="periodic")
        dx, dy, d2x, d2y = cx(t[:, 6], 1), cy(t[:, 6], 1), cx(t[:, 6], 2), cy(t[:, 6], 2)
        tracks[f"{name}_curvature"] = (dx * d2y - dy * d2x) / np.power(dx * dx + dy * dy, 1.5)
        tracks[f"{name}_left"] = np.linalg.norm(p - t[:, 9:11], axis=1)
        tracks[f"{name}_right"] = -np.linalg.norm(p - t[:, 11:13], axis=1)
        tracks[f"{name}_length"] = np.array(t[0, 7] + t[0, 6])
        # the raw table (TrajectoryIndex columns, racing_trajectory.hpp:37-56): input of the track-spline restatement
        tracks[f"{name}_table"] = t
    np.savez_compressed(f"{OUT}/tracks.npz", **tracks)
    for k, v in tracks.items():
        if k.endswith("_length"):
            print(k, float(v))


if __name__ == "__main__":
    main()
