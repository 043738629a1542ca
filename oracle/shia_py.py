"""CPU ORACLE, second opinion -- test infrastructure, NOT product code (parity pinned against the reference's own compiled Fortran, see wmf_oracle.h).

A pure-Python (numpy float32 scalars) transliteration of the core of shia_v1, written from the Fortran source itself
(wmf/modelosv2.f90:257-304 setup, :433-512 vertical tanks, :548-595 hillslope outflow, :599-719 routing, :745-787
records, :797 mean rain, :1278-1293 calc_speed) and *not* from oracle/wmf_oracle_models.c, so that the two restatements
check each other: tests/test_oracle_shia.py requires them to agree to the last bit on small cases.  Sub-models
(sediments, slides, floods, separate_*) are not covered.  Python loops: keep N * T below ~1e5.
"""
import numpy as np

f32 = np.float32


def calc_speed(sm, coef, expo, elem_long, speed, dt, niter):
    """modelosv2.f90:1278-1293; returns (speed, area)"""
    area = f32(0)
    for _ in range(niter):
        area = f32(sm / f32(elem_long + f32(speed * dt)))
        new_speed = f32(coef * f32(area ** expo))
        speed = f32(f32(f32(f32(2) * new_speed) + speed) / f32(3))
    return speed, area


def shia_v1(inp):
    """inp: the dict made by wmf_b200.synth.make_shia_inputs.  Returns dict(q, stoout, hum, mean_rain, hspeed)."""
    N, T = int(inp["nceldas"]), int(inp["n_reg"])
    drena = np.asarray(inp["drena"]).reshape(3, -1, order="F")[0] if np.asarray(inp["drena"]).ndim == 2 else np.asarray(inp["drena"])
    A = lambda k, K: np.asarray(inp[k], dtype=np.float32).reshape(K, N, order="F")
    unit_type = np.asarray(inp["unit_type"]).reshape(-1)
    hill_long, stream_long, elem_area = A("hill_long", 1)[0], A("stream_long", 1)[0], A("elem_area", 1)[0]
    v_coef, h_coef, h_exp = A("v_coef", 4), A("h_coef", 4), A("h_exp", 4)
    control = np.asarray(inp["control"]).reshape(-1)
    control_h = np.asarray(inp["control_h"]).reshape(-1)
    calib = np.asarray(inp["calib"], dtype=np.float32)
    dt = f32(inp["dt"])
    speed_type = [int(v) for v in np.asarray(inp.get("speed_type", [1, 1, 1])).reshape(-1)] + [2]
    niter = int(inp.get("calc_niter", 5))
    retorno_gr, retorno_aq = float(inp.get("retorno_gr", 0)), float(inp.get("retorno_aq", 0))
    evp = np.asarray(inp["evpserie"], dtype=np.float32).reshape(-1)
    rain_tab = np.asarray(inp["rain"])
    pos = np.asarray(inp["pos_evento"]).reshape(-1)
    n_cont = int(np.count_nonzero(control)) + 1
    n_conth = max(1, int(np.count_nonzero(control_h)))
    # :267-304
    m3 = (elem_area / f32(1000.0)).astype(np.float32)
    Sto = (A("storage", 5) + f32(inp.get("storage_constant", 0.0))).astype(np.float32)
    vspeed = np.zeros((4, N), np.float32)
    hspeed = np.zeros((4, N), np.float32)
    for i in range(4):
        vspeed[i] = (v_coef[i] * calib[i]) * dt
        if speed_type[i] == 1:
            hspeed[i] = h_coef[i] * calib[i + 4]
    H = np.stack([A("max_capilar", 1)[0] * calib[8], A("max_gravita", 1)[0] * calib[9], A("max_aquifer", 1)[0] * calib[10]]).astype(np.float32)
    Q = np.zeros((n_cont, T), np.float32)
    Hum = np.zeros((n_conth, T), np.float32)
    sep = int(inp.get("separate_fluxes", 0)) == 1
    Fluxes = np.zeros((3, N + 1), np.float32)        # :340-344 (+1: the outlet's receiver)
    Qsep = np.zeros((n_cont, 3, T), np.float32)
    mean_rain = np.zeros(T, np.float32)
    hflux = [f32(0)] * 4
    for t in range(T):
        control_cont, controlh_cont = 1, 0          # 0-based rows: row 0 is the outlet
        if pos[t] == 1:
            Rain = np.zeros(N, np.float32)
        else:
            Rain = (rain_tab[pos[t] - 1].astype(np.float32) / f32(1000.0)).astype(np.float32)
        rain_sum = f32(0)
        for c in range(N):
            drenaid = N - int(drena[c])              # 0-based index of the cell it drains to (N when drena == 0)
            R = Rain[c]
            rain_sum = f32(rain_sum + R)
            # vertical tanks :478-512
            vflux = [f32(0)] * 4
            vflux[0] = max(f32(0), f32(f32(R - H[0, c]) + Sto[0, c]))
            Sto[0, c] = f32(f32(Sto[0, c] + R) - vflux[0])
            e = f32(f32(evp[t] * vspeed[0, c]) * f32(f32(Sto[0, c] / H[0, c]) ** f32(0.6)))
            evp_loss = min(e, Sto[0, c])
            Sto[0, c] = f32(Sto[0, c] - evp_loss)
            for i in range(3):
                vflux[i + 1] = min(vflux[i], vspeed[i + 1, c])
                Sto[i + 1, c] = f32(f32(Sto[i + 1, c] + vflux[i]) - vflux[i + 1])
            if retorno_aq > 0:
                ret = max(f32(0), f32(Sto[3, c] - H[2, c]))
                Sto[2, c] = f32(Sto[2, c] + ret)
                Sto[3, c] = f32(Sto[3, c] - ret)
            if retorno_gr > 0:
                ret = max(f32(0), f32(Sto[2, c] - H[1, c]))
                Sto[1, c] = f32(Sto[1, c] + ret)
                Sto[2, c] = f32(Sto[2, c] - ret)
            # hillslope outflow :553-595
            for i in range(3):
                if speed_type[i] == 1:
                    hflux[i] = f32(f32(f32(1) - f32(hill_long[c] / f32(f32(hspeed[i, c] * dt) + hill_long[c]))) * Sto[i + 1, c])
                else:
                    hspeed[i, c], area = calc_speed(f32(Sto[i + 1, c] * m3[c]), f32(h_coef[i, c] * calib[i + 4]), h_exp[i, c],
                                                    hill_long[c], hspeed[i, c], dt, niter)
                    hflux[i] = min(f32(f32(f32(area * hspeed[i, c]) * dt) / m3[c]), Sto[i + 1, c])
                Sto[i + 1, c] = f32(Sto[i + 1, c] - hflux[i])
            # routing :599-719
            ut = int(unit_type[c])
            if ut == 1:
                if drena[c] != 0:
                    for i in range(3):
                        Sto[i + 1, drenaid] = f32(Sto[i + 1, drenaid] + hflux[i])
                else:
                    Q[0, t] = f32(Q[0, t] + f32(f32(f32(hflux[0] + hflux[1]) + hflux[2]) * m3[c]))
            else:
                if drena[c] != 0:                    # the Fortran writes StoOut(4, N_cel+1) here for the outlet: dropped
                    Sto[3, drenaid] = f32(Sto[3, drenaid] + f32(hflux[2] * f32(3 - ut)))
                Sto[4, c] = f32(f32(Sto[4, c] + f32(hflux[0] + hflux[1])) + f32(hflux[2] * f32(ut - 2)))
                hspeed[3, c], area = calc_speed(f32(Sto[4, c] * m3[c]), f32(h_coef[3, c] * calib[7]), h_exp[3, c],
                                                stream_long[c], hspeed[3, c], dt, niter)
                hflux[3] = min(f32(f32(f32(area * hspeed[3, c]) * dt) / m3[c]), Sto[4, c])
                if sep:                              # :658-665
                    for i in range(3):
                        Fluxes[i, c] = f32(Fluxes[i, c] + hflux[i])
                    if Sto[4, c] > 0:
                        r = f32(hflux[3] / Sto[4, c])
                        for i in range(3):
                            Fluxes[i, c] = f32(Fluxes[i, c] - f32(Fluxes[i, c] * r))
                    else:
                        Fluxes[:, c] = 0
                Sto[4, c] = f32(Sto[4, c] - hflux[3])
                if drena[c] != 0:
                    Sto[4, drenaid] = f32(Sto[4, drenaid] + hflux[3])
                    if sep:                          # :674-680
                        if Sto[4, c] > 0:
                            r = f32(hflux[3] / Sto[4, c])
                            for i in range(3):
                                Fluxes[i, drenaid] = f32(Fluxes[i, drenaid] + f32(Fluxes[i, c] * r))
                        else:
                            Fluxes[:, drenaid] = 0
                else:
                    Q[0, t] = f32(f32(hflux[3] * m3[c]) / dt)
                    if sep:                          # :697-704
                        for i in range(3):
                            Qsep[0, i, t] = f32(f32(f32(Fluxes[i, c] * f32(hflux[3] / Sto[4, c])) * m3[c]) / dt) if Sto[4, c] > 0 else f32(0)
            # records :749-787 (hflux(4) is whatever the last channel cell left there -- the reference's quirk)
            if control[c] != 0:
                if control_cont < n_cont:
                    Q[control_cont, t] = f32(f32(hflux[3] * m3[c]) / dt)
                    if sep and ut > 1:               # :752-759
                        for i in range(3):
                            Qsep[control_cont, i, t] = f32(f32(f32(Fluxes[i, c] * f32(hflux[3] / Sto[4, c])) * m3[c]) / dt) if Sto[4, c] > 0 else f32(0)
                control_cont += 1
            if control_h[c] != 0:
                if controlh_cont < n_conth:
                    Hum[controlh_cont, t] = f32(Sto[0, c] + Sto[2, c])
                controlh_cont += 1
        mean_rain[t] = f32(rain_sum / f32(N))
    return dict(q=Q, stoout=Sto, hum=Hum, mean_rain=mean_rain, hspeed=hspeed, qseparated=Qsep)
