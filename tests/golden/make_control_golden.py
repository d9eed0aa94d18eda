"""Golden vectors for the node loop after the solve, from the reference's own statements.

scripts/point_follower_local_planner.py imports rclpy and its loop body lives inside main(), so the statements of
main()'s `while True:` loop that follow `u = mpc.perform_mpc(...)` (the acceleration limiter :196-205 and the goal-reached
logic :207-231) are cut out of the unmodified file with `ast` and executed as they are, with recording stand-ins for
cmd_vel_publisher / robot_controller.  The command a robot executes in a step is the LAST one published in that step.
Output: tests/golden/control_golden.npz — single steps from random states and multi-step sequences."""
import ast
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/ros2_mpc/scripts/point_follower_local_planner.py"


def loop_tail():
    tree = ast.parse(open(SRC).read())
    main = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "main")
    loop = next(n for n in ast.walk(main) if isinstance(n, ast.While))
    idx = next(i for i, st in enumerate(loop.body)
               if isinstance(st, ast.Assign) and isinstance(st.value, ast.Call)
               and getattr(st.value.func, "attr", "") == "perform_mpc")
    tail = loop.body[idx + 1:]
    assert len(tail) == 2 and all(isinstance(t, ast.If) for t in tail), "the reference's loop changed"
    return compile(ast.Module(body=tail, type_ignores=[]), SRC, "exec")


class _Pub:
    def __init__(self):
        self.cmds = []

    def publish_cmd(self, v, w):
        self.cmds.append((float(v), float(w)))


class _Log:
    def get_logger(self):
        return self

    def info(self, *_):
        pass


def step(code, u, u_last, x0, goal, flag, goal_threshold):
    pub = _Pub()
    ns = dict(np=np, cmd_vel_publisher=pub, robot_controller=_Log(), GOAL_FLAG=bool(flag), u=u, u_last=u_last, x0=x0,
              goal=goal, goal_threshold=goal_threshold)
    exec(code, ns)
    cmd = pub.cmds[-1] if pub.cmds else (np.nan, np.nan)
    return np.array(cmd), bool(ns["GOAL_FLAG"]), np.asarray(ns["u_last"], dtype=np.float64)


def main():
    code = loop_tail()
    rng = np.random.default_rng(31)
    thr = 0.2
    # ---- single steps ----
    S = 3000
    u = np.round(rng.uniform(-0.2, 0.2, (S, 2)), 3)
    ul = np.round(rng.uniform(-0.2, 0.2, (S, 2)), 3)
    near = rng.random(S) < 0.5
    ul[near] = u[near] + rng.uniform(-0.035, 0.035, (near.sum(), 2))      # around the 0.03 limiter threshold
    x0 = np.c_[np.round(rng.uniform(-2, 2, (S, 2)), 2), rng.uniform(0, 6.28, S)]
    goal = np.zeros((S, 5))
    goal[:, :2] = x0[:, :2] + rng.uniform(-0.3, 0.3, (S, 2))              # around the goal threshold
    on = rng.random(S) < 0.2
    d = rng.uniform(0, 2 * np.pi, on.sum())
    goal[on, :2] = x0[on, :2] + 0.2 * np.c_[np.cos(d), np.sin(d)]          # (almost) exactly on the threshold circle
    flag = rng.random(S) < 0.4
    cmd = np.empty((S, 2)); fo = np.empty(S, bool); ulo = np.empty((S, 2))
    for i in range(S):
        cmd[i], fo[i], ulo[i] = step(code, u[i].copy(), ul[i].copy(), x0[i], goal[i], flag[i], thr)
    out = dict(goal_threshold=np.array(thr), s_u=u, s_u_last=ul, s_x0=x0, s_goal=goal, s_flag=flag, s_cmd=cmd, s_flag_out=fo,
               s_u_last_out=ulo)
    # ---- sequences: u_last and GOAL_FLAG carried from step to step, as the node does (u_last starts as np.array([0, 0])) ----
    Q, T = 64, 40
    us = np.round(np.cumsum(rng.normal(0, 0.02, (Q, T, 2)), axis=1), 3)
    xs = np.c_[np.round(rng.uniform(-1, 1, (Q, 2)), 2), np.zeros(Q)]
    gs = np.zeros((Q, 5)); gs[:, :2] = xs[:, :2] + rng.uniform(-0.6, 0.6, (Q, 2))
    path = np.linspace(0, 1, T)[None, :, None] * np.where(np.arange(Q) % 2 == 0, 1.0, 0.35)[:, None, None]  # half stop short
    xt = xs[:, None, :2] * (1 - path) + gs[:, None, :2] * path                 # the robot approaches its goal
    xt = np.round(xt + rng.normal(0, 0.01, xt.shape), 2)
    cmds = np.empty((Q, T, 2)); flags = np.empty((Q, T), bool); uls = np.empty((Q, T, 2))
    for q in range(Q):
        ul_q, fl_q = np.array([0, 0]), False
        for t in range(T):
            x0_q = np.array([xt[q, t, 0], xt[q, t, 1], 0.0])
            cmds[q, t], fl_q, ul_q = step(code, us[q, t].copy(), ul_q, x0_q, gs[q], fl_q, thr)
            flags[q, t] = fl_q; uls[q, t] = ul_q
    out.update(q_u=us, q_x0=xt, q_goal=gs, q_cmd=cmds, q_flag=flags, q_u_last=uls)
    np.savez_compressed(os.path.join(HERE, "control_golden.npz"), **out)
    print("control_golden: limiter active in", float((np.linalg.norm(u - ul, axis=1) > 0.03).mean()), "of the single steps;",
          "goal flag set at the end of", int(flags[:, -1].sum()), "of", Q, "sequences")


if __name__ == "__main__":
    main()
