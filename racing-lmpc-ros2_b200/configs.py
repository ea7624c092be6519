"""Configuration ground truth of the reference, restated as plain dicts.

Values are the reference's shipped YAMLs (paths under
/root/reference/src/launch/racing_lmpc_launch/param/):
  barc/barc_base.param.yaml:8-66,146-151, barc/barc_single_track.param.yaml:4-11,
  iac_car/iac_car_base.param.yaml, iac_car/iac_car_single_track.param.yaml,
  racing_mpc/barc_lmpc.param.yaml, racing_mpc/barc_tracking_mpc.param.yaml,
  racing_mpc/iac_car_tracking_mpc.param.yaml, racing_mpc/iac_car_lmpc.param.yaml,
  hawaii_gokart/hawaii_gokart_base.param.yaml, hawaii_gokart/hawaii_gokart_single_track.param.yaml,
  racing_mpc/hawaii_kart_tracking_mpc.param.yaml.
The horizon N is a parameter (racing_mpc_config.hpp:47); BASELINE.json's configs override the
YAML values (20 / 40).  Keys mirror include/lmpc_b200.h's lmpc_vehicle_params / lmpc_mpc_config.
"""
import math

INF = math.inf

BARC_VEHICLE = dict(
    mass=2.2187, moi=0.02723, wheel_base=0.324, cg_ratio=0.5, cg_height=0.07, fr=0.012,
    chassis_b=0.281, kd=0.0, kb=0.5, air_density=1.2, frontal_area=1.0, drag_coeff=0.0,
    cl_f=0.0, cl_r=0.0, mu=0.9, Bf=5.0, Cf=2.28, Br=5.0, Cr=2.28,
    Fd_max=15.0, Fb_max=-15.0, Td=0.1, Tb=0.1, max_steer=0.314159, max_steer_rate=10.0,
    integrator=0)

IAC_VEHICLE = dict(
    mass=811.9303, moi=700.0, wheel_base=2.9718, cg_ratio=0.45, cg_height=0.35, fr=0.012,
    chassis_b=2.0, kd=0.0, kb=0.54, air_density=1.2, frontal_area=1.0, drag_coeff=1.0,
    cl_f=1.0, cl_r=1.0, mu=1.3, Bf=11.0, Cf=1.7, Br=11.0, Cr=1.7,
    Fd_max=10000.0, Fb_max=-20000.0, Td=0.1, Tb=0.1, max_steer=0.314159, max_steer_rate=0.66,
    integrator=0)

# hawaii_gokart_base.param.yaml:43-66 (chassis, aero, front brake bias 0.0: the kart brakes on the rear axle only),
# hawaii_gokart_single_track.param.yaml:4-10
HAWAII_KART_VEHICLE = dict(
    mass=180.0, moi=180.0, wheel_base=1.05, cg_ratio=0.45, cg_height=0.25, fr=0.012,
    chassis_b=1.0, kd=0.0, kb=0.0, air_density=1.2041, frontal_area=0.4, drag_coeff=0.8,
    cl_f=0.0, cl_r=0.0, mu=1.5, Bf=14.15, Cf=1.77, Br=14.15, Cr=1.77,
    Fd_max=1000.0, Fb_max=-2000.0, Td=0.1, Tb=0.5, max_steer=0.314159, max_steer_rate=0.5,
    integrator=0)


def barc_lmpc_config(N=20):
    """racing_mpc/barc_lmpc.param.yaml (learning: true)."""
    return dict(
        N=N, learning=1, margin=0.1, q_contour=1.0, q_heading=1.0, q_vel=0.2, q_vy=0.001,
        q_vyaw=0.001, q_boundary=1000.0, R=[0.1, 0.0, 0.0, 0.1], R_d=[0.1, 0.0, 0.0, 0.1],
        x_max=[INF, INF, INF, 3.0, 1.0, 3.0], x_min=[-INF, -INF, -INF, 0.1, -1.0, -3.0],
        u_max=[0.01, 0.33], u_min=[-0.01, -0.33],
        convex_hull_slack=[40.0, 40.0, 4.0, 40.0, 40.0, 4.0],
        num_ss_pts=96, num_ss_pts_per_lap=32, max_lap_stored=3, max_iter=30, tol=1e-7)


def barc_tracking_config(N=20):
    """racing_mpc/barc_tracking_mpc.param.yaml (learning: false)."""
    return dict(
        N=N, learning=0, margin=0.1, q_contour=1.0, q_heading=1.0, q_vel=0.2, q_vy=0.001,
        q_vyaw=0.001, q_boundary=20.0, R=[0.01, 0.0, 0.0, 0.01], R_d=[0.01, 0.0, 0.0, 0.01],
        x_max=[INF, INF, INF, 6.0, 1.0, 3.0], x_min=[-INF, -INF, -INF, 0.1, -1.0, -3.0],
        u_max=[0.01, 0.33], u_min=[-0.01, -0.33],
        convex_hull_slack=[20.0, 20.0, 2.0, 20.0, 20.0, 2.0],
        num_ss_pts=96, num_ss_pts_per_lap=32, max_lap_stored=3, max_iter=30, tol=1e-7)


def iac_tracking_config(N=40):
    """racing_mpc/iac_car_tracking_mpc.param.yaml (learning: false)."""
    return dict(
        N=N, learning=0, margin=0.5, q_contour=1.0, q_heading=1.0, q_vel=0.2, q_vy=0.01,
        q_vyaw=0.01, q_boundary=20.0, R=[1e-5, 0.0, 0.0, 1.0], R_d=[1e-4, 0.0, 0.0, 10.0],
        x_max=[INF, INF, INF, 100.0, 15.0, 2.0], x_min=[-INF, -INF, -INF, 3.0, -15.0, -2.0],
        u_max=[5.0, 0.314159], u_min=[-10.0, -0.314159],
        convex_hull_slack=[20.0, 20.0, 2.0, 20.0, 20.0, 2.0],
        num_ss_pts=96, num_ss_pts_per_lap=32, max_lap_stored=3, max_iter=30, tol=1e-7)


def iac_lmpc_config(N=60):
    """racing_mpc/iac_car_lmpc.param.yaml (learning: true, n: 60; launched at dt = 0.1 by sim_putnam_short_lmpc.launch.py:81)."""
    return dict(iac_tracking_config(N), learning=1, R=[1e-4, 0.0, 0.0, 1e-3], R_d=[5e-4, 0.0, 0.0, 1e-1],
                convex_hull_slack=[200.0, 20.0, 2.0, 200.0, 2.0, 20.0])


def hawaii_kart_tracking_config(N=10):
    """racing_mpc/hawaii_kart_tracking_mpc.param.yaml (learning: false, n: 10, margin 0, control weights 1e-12 on the
    longitudinal command; the reference's own solver options there are tol 1e-3 / max_iter 100 for OSQP).  The file is
    stale against the loader: it has no q_vy / q_vyaw, which ros_param_loader.cpp:78-79 requires -- taken as 0 here."""
    return dict(
        N=N, learning=0, margin=0.0, q_contour=0.3, q_heading=0.5, q_vel=0.1, q_vy=0.0,
        q_vyaw=0.0, q_boundary=50.0, R=[1e-12, 0.0, 0.0, 0.1], R_d=[1e-12, 0.0, 0.0, 0.1],
        x_max=[INF, INF, INF, 30.0, 15.0, 2.0], x_min=[-INF, -INF, -INF, 0.1, -15.0, -2.0],
        u_max=[1000.0, 0.314159], u_min=[-2500.0, -0.314159],
        convex_hull_slack=[4000.0, 4000.0, 400.0, 4000.0, 4000.0, 400.0],
        num_ss_pts=48, num_ss_pts_per_lap=18, max_lap_stored=3, max_iter=30, tol=1e-7)


BARC_DT = 0.025           # launch/barc/sim_barc_lmpc.launch.py:81
BARC_TRACK_LENGTH = 17.014223730977069   # 02_barc_center.txt row 1 col 8
PUTNAM_TRACK_LENGTH = 2849.4188497050591  # 10_putnam_optm.txt row 1 col 8
