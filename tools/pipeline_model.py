"""Fluid model of the k_step operand pipeline on one SM (a planning aid, not a measurement).

Each CTA alternates 16 k-slabs of DMMA work (W cycles of the FP64 pipe each) and an epilogue of E cycles that does not
use the pipe.  A slab can start when its TMA load has landed; the load of slab i + S - 1 is issued when slab i - 1 has
been consumed (S stages) and lands L cycles later.  CTAs that are computing at the same time share the pipe equally.
`continuous` keeps prefetching through the epilogue (k_step2); otherwise the pipeline restarts after it (k_step).
usage: python tools/pipeline_model.py
"""
import itertools


def simulate(n_cta, stages, lat, w=2048.0, epi=6000.0, continuous=True, chunks=40, dt=16.0, phase=0.37):
    nsl = 16
    total = chunks * nsl
    st = []
    for c in range(n_cta):
        st.append(dict(done=0, prog=0.0, epi_left=0.0, landed=[None] * (total + stages + 1), t0=c * phase * (nsl * w + epi)))
    t = 0.0
    busy = 0.0

    def issue(s, i, now):
        if i < total and s["landed"][i] is None:
            s["landed"][i] = now + lat
    for s in st:
        for i in range(stages - 1):
            issue(s, i, s["t0"])
    while any(s["done"] < total for s in st):
        ready = []
        for s in st:
            if s["done"] >= total or t < s["t0"]:
                continue
            if s["epi_left"] > 0:
                s["epi_left"] -= dt
                if s["epi_left"] <= 0 and not continuous:            # restart the pipeline after the epilogue
                    for i in range(s["done"], s["done"] + stages - 1):
                        issue(s, i, t)
                continue
            i = s["done"]
            la = s["landed"][i]
            if la is not None and la <= t:
                ready.append(s)
        if ready:
            busy += dt
            share = dt / len(ready)
            for s in ready:
                s["prog"] += share
                if s["prog"] >= w:
                    s["prog"] -= w
                    s["done"] += 1
                    i = s["done"]
                    nxt = i + stages - 2                           # stage of slab i-1 is free again
                    if continuous or (nxt // nsl == (i - 1) // nsl):
                        issue(s, nxt, t)
                    if i % nsl == 0:
                        s["epi_left"] = epi
        t += dt
    return busy / t


if __name__ == "__main__":
    print("lat   2CTA S=3 (k_step)  2CTA S=3 cont (k_step2)  1CTA S=3  1CTA S=7 cont  2CTA S=4 cont")
    for lat in (1000, 2000, 3000, 4000, 6000):
        row = [simulate(2, 3, lat, continuous=False, epi=12000), simulate(2, 3, lat, continuous=True, epi=6000),
               simulate(1, 3, lat, continuous=True, epi=6000), simulate(1, 7, lat, continuous=True, epi=4000),
               simulate(2, 4, lat, continuous=True, epi=6000)]
        print("%5d " % lat + "  ".join("%8.3f" % v for v in row))
