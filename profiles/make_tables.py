#!/usr/bin/env python
"""Prints the results tables of profiles/README.md from the archived bench lines (profiles/r02_bench_*.json).
usage: python profiles/make_tables.py"""
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))

ROWS = [
    ("C1 demo sphere, 1080p (per triangle)", "r02_bench_c1.json", None),
    ("C2 1 M small tri, 1080p", "r02_bench_default.json", None),
    ("C3 50 k large tri, 4K", "r02_bench_c3.json", None),
    ("C3 leg of the default line", "r02_bench_default.json", "c3"),
    ("C4 20 M tri, 16K² (leg of the default line)", "r02_bench_default.json", "c4_bands"),
    ("C5 256 views × 2 M tri", "r02_bench_c5.json", None),
    ("4K, 500 tri (C3 × 0.01)", "r02_bench_c3_scale0.01.json", None),
    ("4K, 2 500 tri (C3 × 0.05)", "r02_bench_c3_scale0.05.json", None),
    ("4K, 10 000 tri (C3 × 0.2)", "r02_bench_c3_scale0.2.json", None),
    ("C2 Phong", "r02_bench_c2_phong.json", None),
    ("C3 Phong", "r02_bench_c3_phong.json", None),
    ("C2 textured", "r02_bench_c2_textured.json", None),
    ("C3 textured", "r02_bench_c3_textured.json", None),
    ("C2 textured + Phong", "r02_bench_c2_textured_phong.json", None),
    ("C3 textured + Phong", "r02_bench_c3_textured_phong.json", None),
]


def load(name):
    p = os.path.join(HERE, name)
    if not os.path.exists(p):
        return None
    lines = [l for l in open(p).read().strip().splitlines() if l.startswith("{")]
    return json.loads(lines[-1]) if lines else None


def main():
    print("| Config | frame (median / best) | throughput | e2e (host buffers) | CPU reference (16 threads) | kernels (ms): set-up / scan / scatter / raster | image = oracle |")
    print("|---|---|---|---|---|---|---|")
    for label, name, leg in ROWS:
        d = load(name)
        if d is None:
            continue
        if leg:
            d = d.get("legs", {}).get(leg)
            if d is None:
                continue
        s = d.get("stage_ms", {})
        e = d.get("e2e") or {}
        cb = d.get("cpu_baseline") or {}
        unit = d["unit"].replace("Mtriangles/s", "Mtri/s").replace("Mpixels/s", "Mpix/s")
        cpu = f"{cb['value']:.1f} {unit}" if cb.get("value") else "—"
        e2e = f"{e['value']:,.0f} {unit} ({e['ms_per_step']:.2f} ms)" if e.get("value") else "—"
        print(f"| {label} | {d['ms_per_step']:.3f} / {d.get('ms_per_step_best', d['ms_per_step']):.3f} ms | {d['value']:,.0f} {unit} | {e2e} | {cpu} | "
              + " / ".join(f"{s.get(k, 0):.3f}" for k in ("setup_kernel", "tile_scan_kernel", "scatter_kernel", "raster_kernel"))
              + f" | {d.get('image_ok')} |")
    d = load("r02_bench_default.json")
    if d:
        r = d["roofline"]
        print("\nC2 roofline:", json.dumps({k: r[k] for k in ("kernel", "achieved", "peak", "frac", "traffic", "kernel_ms")}), "frame", json.dumps(r["frame"]))
        print("C2 issue:", json.dumps(r.get("issue")))
        print("C2 equivalent z-buffer:", json.dumps(r.get("equivalent_zbuffer")))
        c3 = d.get("legs", {}).get("c3", {}).get("roofline", {})
        print("C3 leg roofline:", json.dumps({k: c3.get(k) for k in ("kernel", "achieved", "frac", "traffic", "kernel_ms", "frame")}))
        print("C3 leg issue:", json.dumps(c3.get("issue")))
        print("C3 leg equivalent z-buffer:", json.dumps(c3.get("equivalent_zbuffer")))
        print("clocks:", json.dumps(d.get("clocks")))
        cb = d.get("cpu_baseline", {})
        print("cpu scalar_mt:", json.dumps(cb.get("scalar_mt")))
        for k, v in (cb.get("avx_mt", {}).get("scenes") or {}).items():
            print("cpu avx_mt", k, json.dumps({q: v[q] for q in ("triangles", "ms_median", "value", "gpu_e2e_ms_median", "gpu_e2e_value")}))
    for n in (2, 4, 8):
        d = load(f"r02_bench_default_n{n}.json")
        if d:
            print(f"\nN={n}: C2 {d['ms_per_step']:.3f} ms {d['value']:,.0f} {d['unit']}; with_gather {json.dumps(d.get('with_gather'))}; nccl {json.dumps(d.get('with_gather_nccl'))}; e2e {json.dumps(d.get('e2e'))}")
            c4 = d.get("legs", {}).get("c4_bands")
            if c4:
                print(f"   c4_bands {c4['ms_per_step']:.3f} ms {c4['value']:,.0f} {c4['unit']} image_ok {c4.get('image_ok')} stage {json.dumps(c4.get('stage_ms'))}\n"
                      f"   with_gather {json.dumps(c4.get('with_gather'))}\n   nccl {json.dumps(c4.get('with_gather_nccl'))}\n   e2e {json.dumps(c4.get('e2e'))}")


if __name__ == "__main__":
    main()
