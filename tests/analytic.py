"""Exact posteriors of the conjugate / finite-state target models (independent analytic oracles,
SURVEY.md §8c "Independent analytic oracles").  numpy only."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden.json")


def golden():
    with open(GOLDEN) as f:
        return json.load(f)


def gaussian_unknown_mean(obs, mu0, s0, s):
    """posterior mean, variance and log evidence of mu ~ N(mu0, s0^2), x_i ~ N(mu, s^2)"""
    obs = np.asarray(obs, float)
    prec = 1 / s0 ** 2 + len(obs) / s ** 2
    var = 1 / prec
    mean = (mu0 / s0 ** 2 + obs.sum() / s ** 2) * var
    # evidence through sequential predictive densities
    m, v, le = mu0, s0 ** 2, 0.0
    for x in obs:
        pv = v + s ** 2
        le += -0.5 * (np.log(2 * np.pi * pv) + (x - m) ** 2 / pv)
        k = v / pv
        m, v = m + k * (x - m), (1 - k) * v
    return mean, var, le


def kalman_smoother(obs):
    """x_t = x_{t-1} + N(0,1), x_0 = 0 given; y_t = x_t + N(0,1).  Returns smoothed means, vars, log evidence."""
    n = len(obs)
    mf, vf, mp, vp = np.zeros(n), np.zeros(n), np.zeros(n), np.zeros(n)
    m, v, le = 0.0, 0.0, 0.0
    for t, y in enumerate(obs):
        mp[t], vp[t] = m, v + 1.0
        s = vp[t] + 1.0
        le += -0.5 * (np.log(2 * np.pi * s) + (y - mp[t]) ** 2 / s)
        k = vp[t] / s
        m, v = mp[t] + k * (y - mp[t]), (1 - k) * vp[t]
        mf[t], vf[t] = m, v
    ms, vs = mf.copy(), vf.copy()
    for t in range(n - 2, -1, -1):
        c = vf[t] / vp[t + 1]
        ms[t] = mf[t] + c * (ms[t + 1] - mp[t + 1])
        vs[t] = vf[t] + c * c * (vs[t + 1] - vp[t + 1])
    return ms, vs, le


HMM_T = np.array([[0.1, 0.5, 0.4], [0.2, 0.2, 0.6], [0.15, 0.15, 0.7]])
HMM_MEANS = np.array([-1.0, 0.0, 1.0])


def hmm_forward_backward(obs):
    """smoothing marginals [n,3] and log evidence of the reference's 3-state HMM"""
    obs = np.asarray(obs, float)
    n = len(obs)
    T = HMM_T / HMM_T.sum(1, keepdims=True)
    lik = np.exp(-0.5 * (obs[:, None] - HMM_MEANS[None, :]) ** 2) / np.sqrt(2 * np.pi)
    alpha = np.zeros((n, 3))
    c = np.zeros(n)
    a = np.full(3, 1 / 3) * lik[0]
    c[0] = a.sum()
    alpha[0] = a / c[0]
    for t in range(1, n):
        a = (alpha[t - 1] @ T) * lik[t]
        c[t] = a.sum()
        alpha[t] = a / c[t]
    beta = np.ones((n, 3))
    for t in range(n - 2, -1, -1):
        beta[t] = (T @ (lik[t + 1] * beta[t + 1])) / c[t + 1]
    g = alpha * beta
    return g / g.sum(1, keepdims=True), np.log(c).sum()
