#!/usr/bin/env python
"""Writes tests/golden/reference_kats.json.

The Julia reference cannot run here (no julia binary), so these golden vectors
are the *hand-derived analytic fixtures held by the reference's own tests*,
transcribed literal for literal (Julia `a / b` on integer literals is float
division, same as Python).  Every entry cites the reference test file:line it
was copied from (paths relative to /root/reference/test/DerivativeOperators/).
Run:  python tests/golden/make_golden.py
"""
import json
import os

import numpy as np

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_kats.json")


def Z(r, c):
    return np.zeros((r, c))


# ---- derivative_operators_interface.jl:8-40 -------------------------------------------------
def fourth_deriv_approx_stencil(N):
    A = Z(N, N + 2)
    r1 = [3.5, -56 / 3, 42.5, -54.0, 251 / 6, -20.0, 5.5, -2 / 3]
    r2 = [2 / 3, -11 / 6, 0.0, 31 / 6, -22 / 3, 4.5, -4 / 3, 1 / 6]
    A[0, 0:8] = r1
    A[1, 0:8] = r2
    A[N - 2, N - 6:] = r2[::-1]
    A[N - 1, N - 6:] = r1[::-1]
    for i in range(3, N - 1):               # i in 3:(N-2); A[i,(i-2):(i+4)]
        A[i - 1, i - 3:i + 4] = [-1 / 6, 2.0, -13 / 2, 28 / 3, -13 / 2, 2.0, -1 / 6]
    return A


def second_deriv_fourth_approx_stencil(N):
    A = Z(N, N + 2)
    A[0, 0:6] = [5 / 6, -15 / 12, -1 / 3, 7 / 6, -6 / 12, 5 / 60]
    A[N - 1, N - 4:] = [1 / 12, -6 / 12, 14 / 12, -4 / 12, -15 / 12, 10 / 12]
    for i in range(2, N):                   # i in 2:(N-1); A[i,(i-1):(i+3)]
        A[i - 1, i - 2:i + 3] = [-1 / 12, 4 / 3, -5 / 2, 4 / 3, -1 / 12]
    return A


def second_derivative_stencil(N):
    A = Z(N, N + 2)
    for i in range(1, N + 1):
        for j in range(1, N + 3):
            if j - i == 0 or j - i == 2:
                A[i - 1, j - 1] = 1
            if j - i == 1:
                A[i - 1, j - 1] = -2
    return A


# ---- derivative_operators_interface.jl:48-100 (grid [0,.08,.1,.15,.19,.26,.29]) -------------
def analyticCtrOneTwoIrr():
    A = Z(5, 7)
    A[0, 0:3] = [-5.0 / 2.0, -75.0 / 2.0, 40.0]
    A[1, 1:4] = [-250.0 / 7.0, 30.0, 40.0 / 7.0]
    A[2, 2:5] = [-80.0 / 9.0, -5.0, 125.0 / 9.0]
    A[3, 3:6] = [-1225.0 / 77.0, 825.0 / 77.0, 400.0 / 77.0]
    A[4, 4:7] = [-30.0 / 7.0, -400.0 / 21.0, 70.0 / 3.0]
    return A


def analyticCtrTwoTwoIrr():
    A = Z(5, 7)
    A[0, 0:3] = [250.0, -1250.0, 1000.0]
    A[1, 1:4] = [10000.0 / 7.0, -2000.0, 4000.0 / 7.0]
    A[2, 2:5] = [4000.0 / 9.0, -1000.0, 5000.0 / 9.0]
    A[3, 3:6] = [454.0 + 42.0 / 77.0, -5000.0 / 7.0, 20000.0 / 77.0]
    A[4, 4:7] = [6000.0 / 21.0, -20000.0 / 21.0, 2000.0 / 3.0]
    return A


def analyticCtrTwoFourIrr():
    A = Z(3, 7)
    A[0, 0:5] = [14.0 + 12012.0 / 13167.0, 1542.0 + 2736.0 / 13167.0, -2288.0 - 11704.0 / 13167.0,
                 838.0 + 1254.0 / 13167.0, -106.0 - 4298.0 / 13167.0]
    A[1, 1:6] = [-223.0 - 461.0 / 693.0, 847.0 + 154.0 / 693.0, -1311.0 - 477.0 / 693.0,
                 699.0 + 593.0 / 693.0, -11.0 - 502.0 / 693.0]
    A[2, 2:7] = [2.0 + 12166.0 / 13167.0, 538.0 + 12654.0 / 13167.0, -912.0 - 9196.0 / 13167.0,
                 508.0 + 8664.0 / 13167.0, -137.0 - 11121.0 / 13167.0]
    return A


def analyticCtrFourTwoIrr():
    A = Z(3, 7)
    A[0, 0:5] = [462000000.0 / 4389.0, -8550000000.0 / 4389.0, 11704000000.0 / 4389.0,
                 -5016000000.0 / 4389.0, 1400000000.0 / 4389.0]
    A[1, 1:6] = [200000000.0 / 231.0, -385000000.0 / 231.0, 360000000.0 / 231.0,
                 -200000000.0 / 231.0, 25000000.0 / 231.0]
    A[2, 2:7] = [770000000.0 / 4389.0, -3420000000.0 / 4389.0, 4180000000.0 / 4389.0,
                 -2850000000.0 / 4389.0, 1320000000 / 4389.0]
    return A


# ---- upwind_operators_interface.jl:12-86 ------------------------------------------------------
def analyticOneOnePos():
    A = Z(5, 7)
    for i in range(1, 6):
        A[i - 1, i:i + 2] = [-1, 1]
    return A


def analyticOneOneNeg():
    A = Z(5, 7)
    for i in range(1, 6):
        A[i - 1, i - 1:i + 1] = [-1, 1]
    return A


def analyticOneTwoPos():
    A = Z(5, 7)
    for i in range(1, 5):
        A[i - 1, i:i + 3] = [-3 / 2, 2, -1 / 2]
    A[4, 4:7] = [-1 / 2, 0, 1 / 2]
    return A


def analyticOneTwoNeg():
    A = Z(5, 7)
    A[0, 0:3] = [-1 / 2, 0, 1 / 2]
    for i in range(2, 6):
        A[i - 1, i - 2:i + 1] = [1 / 2, -2, 3 / 2]
    return A


def analyticTwoTwoPos():
    A = Z(5, 7)
    for i in range(1, 4):
        A[i - 1, i:i + 4] = [2, -5, 4, -1]
    A[3, 3:7] = [1, -2, 1, 0]
    A[4, 3:7] = [0, 1, -2, 1]
    return A


def analyticTwoTwoNeg():
    A = Z(5, 7)
    A[0, 0:4] = [1, -2, 1, 0]
    A[1, 0:4] = [0, 1, -2, 1]
    for i in range(3, 6):
        A[i - 1, i - 3:i + 1] = [-1, 4, -5, 2]
    return A


def analyticTwoThreePos():
    A = Z(7, 9)
    for i in range(1, 5):
        A[i - 1, i:i + 5] = [35 / 12, -104 / 12, 114 / 12, -56 / 12, 11 / 12]
    A[4, 4:9] = [11 / 12, -20 / 12, 6 / 12, 4 / 12, -1 / 12]
    A[5, 4:9] = [-1 / 12, 16 / 12, -30 / 12, 16 / 12, -1 / 12]
    A[6, 4:9] = [-1 / 12, 4 / 12, 6 / 12, -20 / 12, 11 / 12]
    return A


def analyticTwoThreeNeg():
    A = Z(7, 9)
    A[0, 0:5] = [11 / 12, -20 / 12, 6 / 12, 4 / 12, -1 / 12]
    A[1, 0:5] = [-1 / 12, 16 / 12, -30 / 12, 16 / 12, -1 / 12]
    A[2, 0:5] = [-1 / 12, 4 / 12, 6 / 12, -20 / 12, 11 / 12]
    for i in range(4, 8):
        A[i - 1, i - 4:i + 1] = [11 / 12, -56 / 12, 114 / 12, -104 / 12, 35 / 12]
    return A


# ---- upwind_operators_interface.jl:92-150 (irregular grid) -----------------------------------
def analyticOneOnePosIrr():
    A = Z(5, 7)
    A[0, 1:3] = [-50, 50]
    A[1, 2:4] = [-20, 20]
    A[2, 3:5] = [-25, 25]
    A[3, 4:6] = [-100 / 7, 100 / 7]
    A[4, 5:7] = [-100 / 3, 100 / 3]
    return A


def analyticOneOneNegIrr():
    A = Z(5, 7)
    A[0, 0:2] = [-25 / 2, 25 / 2]
    A[1, 1:3] = [-50, 50]
    A[2, 2:4] = [-20, 20]
    A[3, 3:5] = [-25, 25]
    A[4, 4:6] = [-100 / 7, 100 / 7]
    return A


def analyticOneTwoPosIrr():
    A = Z(5, 7)
    A[0, 1:4] = [-450 / 7, 490 / 7, -40 / 7]
    A[1, 2:5] = [-280 / 9, 405 / 9, -125 / 9]
    A[2, 3:6] = [-2625 / 77, 3025 / 77, -400 / 77]
    A[3, 4:7] = [-510 / 21, 1000 / 21, -490 / 21]
    A[4, 4:7] = [-90 / 21, -400 / 21, 490 / 21]
    return A


def analyticOneTwoNegIrr():
    A = Z(5, 7)
    A[0, 0:3] = [-5 / 2, -75 / 2, 80 / 2]
    A[1, 0:3] = [5 / 2, -125 / 2, 120 / 2]
    A[2, 1:4] = [250 / 7, -490 / 7, 240 / 7]
    A[3, 2:5] = [80 / 9, -405 / 9, 325 / 9]
    A[4, 3:6] = [1225 / 77, -3025 / 77, 1800 / 77]
    return A


def analyticTwoTwoPosIrr():
    A = Z(5, 7)
    A[0, 1:5] = [200000 / 77, -308000 / 77, 143000 / 77, -35000 / 77]
    A[1, 2:6] = [27500 / 33, -75000 / 33, 55000 / 33, -7500 / 33]
    A[2, 3:7] = [72500 / 77, -137500 / 77, 120000 / 77, -55000 / 77]
    A[3, 3:7] = [42500 / 77, -71500 / 77, 40000 / 77, -11000 / 77]
    A[4, 3:7] = [-10000 / 77, 44000 / 77, -100000 / 77, 66000 / 77]
    return A


def analyticTwoTwoNegIrr():
    A = Z(5, 7)
    A[0, 0:4] = [1050 / 7, -1250 / 7, -1400 / 7, 1600 / 7]
    A[1, 0:4] = [350 / 7, 6250 / 7, -9800 / 7, 3200 / 7]
    A[2, 0:4] = [-1400 / 7, 25000 / 7, -30800 / 7, 7200 / 7]
    A[3, 1:5] = [-390000 / 231, 770000 / 231, -660000 / 231, 280000 / 231]
    A[4, 2:6] = [-38500 / 77, 161000 / 77, -165000 / 77, 42500 / 77]
    return A


IRR_DX = [0.08, 0.02, 0.05, 0.04, 0.07, 0.03]


def mat(cite, kind, d, a, dx, n, coeff, matrix, rows=None, rtol=None, offside=0):
    e = dict(cite=cite, kind=kind, d=d, a=a, dx=dx, n=n, coeff=coeff, offside=offside,
             matrix=np.asarray(matrix).tolist())
    if rows is not None:
        e["rows"] = rows          # 0-based [start, stop) rows of the operator matrix the fixture covers
    if rtol is not None:
        e["rtol"] = rtol
    return e


def main():
    G = {"_about": "Golden vectors transcribed from the reference's own tests; see make_golden.py"}

    # full operator matrices N x (N+2): compare with convert_by_multiplication (mul! on unit vectors)
    G["operator_matrices"] = [
        mat("derivative_operators_interface.jl:8-21,121-127", "centered", 4, 4, 1.0, 20, 1.0, fourth_deriv_approx_stencil(20)),
        mat("derivative_operators_interface.jl:23-31,134-138", "centered", 2, 4, 1.0, 20, 1.0, second_deriv_fourth_approx_stencil(20)),
        mat("derivative_operators_interface.jl:33-40,285-291", "centered", 2, 2, 1.0, 10, 1.0, second_derivative_stencil(10)),
        mat("convolutions.jl(test):45-81", "centered", 4, 4, 1.0, 20, 1.0, fourth_deriv_approx_stencil(20)),
        # non-uniform centered, Array(L) ≈ correct  (:229-281); (2,4) and (4,2) only rows 2:end-1
        mat("derivative_operators_interface.jl:48-56,235-242", "centered", 1, 2, IRR_DX, 5, 1.0, analyticCtrOneTwoIrr()),
        mat("derivative_operators_interface.jl:58-66,247-254", "centered", 2, 2, IRR_DX, 5, 1.0, analyticCtrTwoTwoIrr()),
        mat("derivative_operators_interface.jl:68-80,259-266", "centered", 2, 4, IRR_DX, 5, 1.0, analyticCtrTwoFourIrr(), rows=[1, 4]),
        mat("derivative_operators_interface.jl:82-100,271-278", "centered", 4, 2, IRR_DX, 5, 1.0, analyticCtrFourTwoIrr(), rows=[1, 4]),
        # scalar coefficient  :407-414  Array(3.3*A), A = CenteredDifference(2,2,10.0,3)
        mat("derivative_operators_interface.jl:407-414", "centered", 2, 2, 10.0, 3, 3.3,
            [[0.033, -0.066, 0.033, 0.0, 0.0], [0.0, 0.033, -0.066, 0.033, 0.0], [0.0, 0.0, 0.033, -0.066, 0.033]]),
        # uniform upwind  upwind_operators_interface.jl:152-388 (analyticL = -1*Neg for coefficient -1)
        mat("upwind_operators_interface.jl:12-18,152-169", "upwind", 1, 1, 1.0, 5, 1.0, analyticOneOnePos()),
        mat("upwind_operators_interface.jl:20-26,181-198", "upwind", 1, 1, 1.0, 5, -1.0, -1 * analyticOneOneNeg()),
        mat("upwind_operators_interface.jl:28-35,210-227", "upwind", 1, 2, 1.0, 5, 1.0, analyticOneTwoPos()),
        mat("upwind_operators_interface.jl:37-44,243-260", "upwind", 1, 2, 1.0, 5, -1.0, -1 * analyticOneTwoNeg()),
        mat("upwind_operators_interface.jl:46-54,272-289", "upwind", 2, 2, 1.0, 5, 1.0, analyticTwoTwoPos()),
        mat("upwind_operators_interface.jl:56-64,301-318", "upwind", 2, 2, 1.0, 5, -1.0, -1 * analyticTwoTwoNeg()),
        mat("upwind_operators_interface.jl:66-75,332-349", "upwind", 2, 3, 1.0, 7, 1.0, analyticTwoThreePos()),
        mat("upwind_operators_interface.jl:77-86,361-378", "upwind", 2, 3, 1.0, 7, -1.0, -1 * analyticTwoThreeNeg()),
        # irregular upwind :390-568
        mat("upwind_operators_interface.jl:92-100,390-407", "upwind", 1, 1, IRR_DX, 5, 1.0, analyticOneOnePosIrr()),
        mat("upwind_operators_interface.jl:102-110,420-437", "upwind", 1, 1, IRR_DX, 5, -1.0, -1 * analyticOneOneNegIrr()),
        mat("upwind_operators_interface.jl:112-120,450-467", "upwind", 1, 2, IRR_DX, 5, 1.0, analyticOneTwoPosIrr()),
        mat("upwind_operators_interface.jl:122-130,480-497", "upwind", 1, 2, IRR_DX, 5, -1.0, -1 * analyticOneTwoNegIrr()),
        mat("upwind_operators_interface.jl:132-140,510-527", "upwind", 2, 2, IRR_DX, 5, 1.0, analyticTwoTwoPosIrr()),
        mat("upwind_operators_interface.jl:142-150,540-557", "upwind", 2, 2, IRR_DX, 5, -1.0, -1 * analyticTwoTwoNegIrr()),
        # dx scaling :570-674
        mat("upwind_operators_interface.jl:573-574", "upwind", 1, 2, 0.1, 5, 1.0, 10.0 * analyticOneTwoPos()),
        mat("upwind_operators_interface.jl:600-601", "upwind", 1, 2, 0.1, 5, -1.0, -10.0 * analyticOneTwoNeg()),
        mat("upwind_operators_interface.jl:625-626", "upwind", 2, 2, 0.1, 5, 1.0, 100.0 * analyticTwoTwoPos()),
        mat("upwind_operators_interface.jl:650-651", "upwind", 2, 2, 0.1, 5, -1.0, -100.0 * analyticTwoTwoNeg()),
        # coefficient handling :676-839
        mat("upwind_operators_interface.jl:679-680", "upwind", 2, 2, 1.0, 5, 4.56, 4.56 * analyticTwoTwoPos()),
        mat("upwind_operators_interface.jl:705-706", "upwind", 2, 2, 1.0, 5, -4.56, -4.56 * analyticTwoTwoNeg()),
        mat("upwind_operators_interface.jl:734-735", "upwind", 2, 2, 0.1, 5, 4.56, 4.56 * 100.0 * analyticTwoTwoPos()),
        mat("upwind_operators_interface.jl:760-761", "upwind", 2, 2, 0.1, 5, -4.56, -4.56 * 100.0 * analyticTwoTwoNeg()),
        mat("upwind_operators_interface.jl:789-790", "upwind", 2, 2, IRR_DX, 5, 4.56, 4.56 * analyticTwoTwoPosIrr()),
        mat("upwind_operators_interface.jl:815-816", "upwind", 2, 2, IRR_DX, 5, -4.56, -4.56 * analyticTwoTwoNegIrr()),
    ]

    # interior stencil weights, atol 1e-10 (derivative_operators_interface.jl:146-160, :203-214)
    G["interior_weights"] = {
        "cite": "derivative_operators_interface.jl:149-152 (centered, dx=1), :204-207 (upwind, dx=1, offside=0)",
        "centered": [  # [a][d-1]
            {"a": 2, "d": 1, "w": [-0.5, 0, 0.5]}, {"a": 2, "d": 2, "w": [1.0, -2.0, 1.0]},
            {"a": 2, "d": 3, "w": [-1 / 2, 1.0, 0.0, -1.0, 1 / 2]},
            {"a": 4, "d": 1, "w": [1 / 12, -2 / 3, 0, 2 / 3, -1 / 12]},
            {"a": 4, "d": 2, "w": [-1 / 12, 4 / 3, -5 / 2, 4 / 3, -1 / 12]},
            {"a": 4, "d": 3, "w": [1 / 8, -1.0, 13 / 8, 0.0, -13 / 8, 1.0, -1 / 8]},
        ],
        "upwind": [
            {"a": 1, "d": 1, "w": [-1.0, 1.0]}, {"a": 1, "d": 3, "w": [-1.0, 3.0, -3.0, 1.0]},
            {"a": 2, "d": 1, "w": [-3 / 2, 2.0, -1 / 2]}, {"a": 2, "d": 3, "w": [-5 / 2, 9.0, -12.0, 7.0, -3 / 2]},
        ],
    }

    # vector coefficients  derivative_operators_interface.jl:416-432
    G["vector_coefficients"] = {
        "cite": "derivative_operators_interface.jl:416-432",
        "x": [1.0, 2.0, 3.0, 4.0, 5.0], "c": [1.0, 1.0, 3.0],
        "cC_coefficients": [2.0, 3.0, 12.0],
    }

    # Robin: closed forms for order 1 (robin.jl:13-43 uniform, :46-90 vector dx) with fixed "random" draws,
    # and the 3rd-order ghost values (robin.jl:92-110).
    rng = np.random.default_rng(20240607)
    G["robin_order1"] = {
        "cite": "robin.jl:13-43 (dx::T), :46-90 (dx::Vector = dx .* ones(5i)); expected ghosts are the closed forms at :31-33",
        "cases": [dict(al=float(rng.random()), bl=float(rng.random()), cl=float(rng.random()), dx=float(rng.random()),
                       ar=float(rng.random()), br=float(rng.random()), cr=float(rng.random()),
                       u=rng.random(5 * (i + 1)).tolist()) for i in range(5)],
    }
    G["robin_order3"] = {
        "cite": "robin.jl:92-110",
        "l": [1.0, 6.0, 10.0], "r": [1.0, 6.0, 10.0], "dx": 1.0, "order": 3,
        "u": list(map(float, range(1, 11))), "u0": -4 / 10, "uend": 125 / 12,
        "general_alpha": [-10.0, 1.0, 6.0],
    }

    # L*Q concretized 3x3 matrices (linear part), BasicSDOExamples.jl; dx = 0.25, M = 3
    dx = 0.25
    G["ghost_operator_matrices"] = {
        "cite": "BasicSDOExamples.jl:49-59, :91-101, :137-148, :198-210, :270-277 (x̄ = range(0,1,length=5) -> dx=0.25, M=3)",
        "dx": dx, "M": 3,
        "cases": [
            dict(name="L2bc Neumann0", cite=":58-59", terms=[dict(kind="centered", d=2, a=2, coeff=1.0)],
                 bc=dict(type="neumann0", order=1), matrix=[[-16.0, 16.0, 0.0], [16.0, -32.0, 16.0], [0.0, 16.0, -16.0]]),
            dict(name="L1-bc = -1*Array(Upwind(1,1,dx,3,-1.0)*Q)", cite=":37,:57", scale=-1.0,
                 terms=[dict(kind="upwind", d=1, a=1, coeff=-1.0)], bc=dict(type="neumann0", order=1),
                 matrix=[[0.0, 0.0, 0.0], [-4.0, 4.0, 0.0], [0.0, -4.0, 4.0]]),
            dict(name="L1+bc = Array(Upwind(1,1,dx,3,1.0)*Q)", cite=":85,:99", terms=[dict(kind="upwind", d=1, a=1, coeff=1.0)],
                 bc=dict(type="neumann0", order=1), matrix=[[-4.0, 4.0, 0.0], [0.0, -4.0, 4.0], [0.0, 0.0, 0.0]]),
            dict(name="Lx negative drift mu=-0.1", cite=":41,:55-56",
                 terms=[dict(kind="upwind", d=1, a=1, coeff=1.0, scale=-0.1), dict(kind="centered", d=2, a=2, coeff=1.0, scale=0.1 ** 2 / 2)],
                 bc=dict(type="neumann0", order=1), matrix=[[-0.08, 0.08, 0.0], [0.48, -0.56, 0.08], [0.0, 0.48, -0.48]]),
            dict(name="Lx positive drift mu=0.1", cite=":84,:97-98",
                 terms=[dict(kind="upwind", d=1, a=1, coeff=1.0, scale=0.1), dict(kind="centered", d=2, a=2, coeff=1.0, scale=0.1 ** 2 / 2)],
                 bc=dict(type="neumann0", order=1), matrix=[[-0.48, 0.48, 0.0], [0.08, -0.56, 0.48], [0.0, 0.08, -0.08]]),
            dict(name="L1 state dependent drift mu(x)=-x", cite=":121-128,:145-146",
                 terms=[dict(kind="upwind", d=1, a=1, coeff=[-0.25, -0.5, -0.75])], bc=dict(type="neumann0", order=1),
                 matrix=[[0.0, 0.0, 0.0], [2.0, -2.0, 0.0], [0.0, 3.0, -3.0]]),
            dict(name="absorbing: L2bc with Robin((1,0,S),(0,1,0))", cite=":188-190,:207-208",
                 terms=[dict(kind="centered", d=2, a=2, coeff=1.0)], bc=dict(type="robin", l=[1.0, 0.0, 3.0], r=[0.0, 1.0, 0.0], order=1),
                 matrix=[[-32.0, 16.0, 0.0], [16.0, -32.0, 16.0], [0.0, 16.0, -16.0]]),
            dict(name="absorbing: L1-bc = Array(Upwind(1,1,dx,3,mu)*Q)/mu, mu=-0.1 (params() default, :8,:201)", cite=":182,:198,:206", scale=1 / -0.1,
                 terms=[dict(kind="upwind", d=1, a=1, coeff=-0.1)], bc=dict(type="robin", l=[1.0, 0.0, 3.0], r=[0.0, 1.0, 0.0], order=1),
                 matrix=[[4.0, 0.0, 0.0], [-4.0, 4.0, 0.0], [0.0, -4.0, 4.0]]),
            dict(name="absorbing: Lxbc = Array(L1*Q)[1] + Array(s2/2*L2*Q)[1], mu=-0.1", cite=":193,:204-205",
                 terms=[dict(kind="upwind", d=1, a=1, coeff=-0.1), dict(kind="centered", d=2, a=2, coeff=1.0, scale=0.1 ** 2 / 2)],
                 bc=dict(type="robin", l=[1.0, 0.0, 3.0], r=[0.0, 1.0, 0.0], order=1),
                 matrix=[[-0.56, 0.08, 0.0], [0.48, -0.56, 0.08], [0.0, 0.48, -0.48]]),
            dict(name="KFE with drift (mu=-0.1, params() default): Upwind(1,1,dx,3,-mu)*Q + s2/2*L2*Q, Robin xi=-2mu/s2", cite=":248-264,:274-275",
                 terms=[dict(kind="upwind", d=1, a=1, coeff=0.1), dict(kind="centered", d=2, a=2, coeff=1.0, scale=0.1 ** 2 / 2)],
                 bc=dict(type="robin", l=[-2 * -0.1 / 0.1 ** 2, 1.0, 0.0], r=[-2 * -0.1 / 0.1 ** 2, 1.0, 0.0], order=1),
                 matrix=[[-0.58, 0.48, 0.0], [0.08, -0.56, 0.48], [0.0, 0.08, -0.48]]),
            dict(name="KFE without drift", cite=":265,:276-277", scale=0.1 ** 2 / 2,
                 terms=[dict(kind="centered", d=2, a=2, coeff=1.0)],
                 bc=dict(type="robin", l=[-2 * -0.1 / 0.1 ** 2, 1.0, 0.0], r=[-2 * -0.1 / 0.1 ** 2, 1.0, 0.0], order=1),
                 matrix=[[-0.18, 0.08, 0.0], [0.08, -0.16, 0.08], [0.0, 0.08, -0.1466666666666667]]),
        ],
    }

    # N-D axis application, differentiation_dimension.jl:25-207: separable fields, each pencil equals the
    # 1-D stencil matrix applied to that pencil.
    G["nd_axis"] = {
        "cite": "differentiation_dimension.jl:25-62 (2-D, (4,4), dx=0.1, 22x22, axes 1,2), :64-122 (3-D (2,2) axes 1-3), "
                ":124-184 (3-D (4,4) axes 1-3), :186-207 (7-D, axis 6, cos field); expected pencil = dx^-d * stencil_matrix * pencil",
        "cases": [
            dict(shape=[22, 22], axis=1, d=4, a=4, dx=0.1, n=20, field="sin"),
            dict(shape=[22, 22], axis=2, d=4, a=4, dx=0.1, n=20, field="sin"),
            dict(shape=[22, 22, 22], axis=1, d=2, a=2, dx=0.1, n=20, field="sin"),
            dict(shape=[22, 22, 22], axis=2, d=2, a=2, dx=0.1, n=20, field="sin"),
            dict(shape=[22, 22, 22], axis=3, d=2, a=2, dx=0.1, n=20, field="sin"),
            dict(shape=[22, 22, 22], axis=1, d=4, a=4, dx=0.1, n=20, field="sin"),
            dict(shape=[22, 22, 22], axis=2, d=4, a=4, dx=0.1, n=20, field="sin"),
            dict(shape=[22, 22, 22], axis=3, d=4, a=4, dx=0.1, n=20, field="sin"),
            dict(shape=[5, 5, 5, 5, 5, 32, 5], axis=6, d=4, a=4, dx=0.1, n=30, field="cos"),
        ],
    }

    with open(OUT, "w") as f:
        json.dump(G, f, indent=1)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
