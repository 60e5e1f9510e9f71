#!/usr/bin/env python3
"""Regenerate the committed fixtures under tests/golden/.  Needs /root/reference (build container only):

  Duck.pt.xz                 assets/Duck.glb baked by rayfinder_b200.baker (PtFormat(path) restatement)
  ref_duck_node_counts.npz   bvh-visualizer node-count image, 512x512 + 1280x720 rows, from the REFERENCE's own
                             compiled rayIntersectBvh (oracle/_ref)          [BASELINE.json configs[0]]
  ref_duck_bvh_test.npz      the 64x64 ray grid of the reference's tests/bvh.cpp:34-102 (rayTMax 1000): hit, t, p,
                             nodesVisited from oracle/_ref
  ref_sky_states.npz         sky_state_new outputs of the reference's hw_skymodel.c for a grid of parameters,
                             + sky_state_radiance samples
  ref_cameras.npz            createCamera outputs of the reference's camera.cpp
  oracle_duck_hdr.npz        Oracle B (oracle/oracle.cpp) HDR sums for a small Duck render: a regression pin of
                             the restatement itself (NOT a reference output; radiance parity is unpinned)
  oracle_duck_deferred.npz   Oracle C (deferred lighting + resolve passes) on a small Duck G-buffer, two frames: inputs and
                             outputs, again a regression pin of the restatement (python make_golden.py --deferred-only
                             regenerates just this one)
"""
import lzma
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import _oracle as O  # noqa: E402
import rayfinder_b200 as rf  # noqa: E402
from rayfinder_b200 import baker  # noqa: E402

G = ROOT / "tests" / "golden"
REF_ASSETS = Path("/root/reference/assets")


def main():
    assert O.have_ref(), "build oracle/_ref first (make -C oracle ref)"
    pt = baker.bake(REF_ASSETS / "Duck.glb")
    raw = pt.dumps()
    (G / "Duck.pt.xz").write_bytes(lzma.compress(raw, preset=9))
    nodes, tris = pt.bvh_nodes, O.triangles9(pt)

    out = {}
    for (w, h) in ((512, 512), (1280, 720)):
        cam = rf.camera_to_array(rf.bvh_visualizer_camera(nodes, w, h))
        counts, _ = O.ref_node_counts(nodes, tris, cam, w, h, rf.FLT_MAX)
        out[f"counts_{w}x{h}"] = counts.astype(np.uint16) if counts.max() < 65536 else counts
        out[f"camera_{w}x{h}"] = cam
    np.savez_compressed(G / "ref_duck_node_counts.npz", **out)
    print("node counts 512x512: sum", int(out["counts_512x512"].sum()), "max", int(out["counts_512x512"].max()))

    # tests/bvh.cpp:34-102 — camera from the triangle AABB with 0.8f float constants, aspect 1
    cam = rf.camera_to_array(rf.bvh_visualizer_camera(nodes, 1, 1, float_constants=True))
    rays = np.zeros((64, 64, 6), dtype=np.float32)
    ray = np.zeros(6, dtype=np.float32)
    for i in range(64):
        for j in range(64):
            O.ref().ref_generate_camera_ray(O._ptr(cam), i / 64.0, j / 64.0, O._ptr(ray))
            rays[i, j] = ray
    hit, p_t, visited = O.ref_intersect(nodes, tris, rays, 1000.0)
    np.savez_compressed(G / "ref_duck_bvh_test.npz", camera=cam, rays=rays, hit=hit, p_t=p_t, nodes_visited=visited)
    print("bvh test grid: hits", int(hit.sum()))

    # sky states
    params, states, rad = [], [], []
    for turbidity in (1.0, 1.5, 3.0, 7.25, 10.0):
        for albedo in ((1.0, 1.0, 1.0), (0.0, 0.0, 0.0), (0.2, 0.5, 0.9)):
            for zenith in (0.0, 30.0, 62.5, 89.0):
                elevation = np.float32(0.5) * np.float32(np.pi) - np.float32(rf.degrees_to_radians(zenith))
                st = np.zeros(33, dtype=np.float32)
                alb = np.array(albedo, dtype=np.float32)
                rc = O.ref().ref_sky_state_new(float(elevation), turbidity, O._ptr(alb), O._ptr(st))
                assert rc == 0
                params.append((turbidity, *albedo, zenith))
                states.append(st)
                rad.append([O.ref().ref_sky_state_radiance(O._ptr(st), th, ga, ch)
                            for th in (0.1, 0.7, 1.5) for ga in (0.2, 1.0, 2.5) for ch in range(3)])
    np.savez_compressed(G / "ref_sky_states.npz", params=np.array(params, dtype=np.float32), states=np.array(states),
                        radiance=np.array(rad, dtype=np.float32))

    cams_in, cams_out = [], []
    rng = np.random.default_rng(7)
    for _ in range(32):
        o = rng.uniform(-5, 5, 3).astype(np.float32)
        la = rng.uniform(-5, 5, 3).astype(np.float32)
        ap, fd = np.float32(rng.uniform(0, 0.5)), np.float32(rng.uniform(0.5, 20))
        vf, asp = np.float32(rng.uniform(20, 110)), np.float32(rng.uniform(0.5, 2.5))
        c = np.zeros(19, dtype=np.float32)
        O.ref().ref_create_camera(O._ptr(o), O._ptr(la), float(ap), float(fd), float(vf), float(asp), O._ptr(c))
        cams_in.append(np.concatenate([o, la, [ap, fd, vf, asp]]).astype(np.float32))
        cams_out.append(c)
    np.savez_compressed(G / "ref_cameras.npz", inputs=np.array(cams_in), cameras=np.array(cams_out))

    # Oracle B regression pin
    w, h, spp, bounces = 96, 64, 2, 4
    cam = rf.camera_to_array(rf.bvh_visualizer_camera(nodes, w, h))
    sky = rf.sky_state(rf.Sky())
    orc = O.OracleRenderer(pt, w, h, cam, sky, spp, bounces, threads=1)
    orc.render()
    orc.render()
    np.savez_compressed(G / "oracle_duck_hdr.npz", image=orc.image, camera=cam, sky=sky, spp=spp, bounces=bounces,
                        counters=orc.counters, path_lengths=orc.path_lengths)
    print("oracle duck hdr: mean", float(orc.image[..., :3].mean()), orc.stats())


def deferred_pin(pt):
    from test_gpu_deferred import make_gbuffer

    w, h = 64, 48
    lo, hi = pt.bvh_nodes["aabb_min"][0].astype(np.float64), pt.bvh_nodes["aabb_max"][0].astype(np.float64)
    centre = 0.5 * (lo + hi)
    eye = centre + np.array([0.9, 0.5, 1.1]) * (hi - lo).max()
    inv, albedo, normal, depth = make_gbuffer(pt, w, h, eye, centre, seed=3)
    sky = rf.sky_state(rf.Sky(turbidity=2.0, sun_zenith_degrees=35.0, sun_azimuth_degrees=20.0))
    orc = O.OracleDeferredLighting(pt, sky, threads=1)
    out = {"inv": inv, "eye": eye.astype(np.float32), "albedo": albedo[..., :3].astype(np.float32), "normal": normal[..., :3].astype(np.float32),
           "depth": depth, "sky": sky}
    for frame in (0, 5):
        orc.render(inv, eye.astype(np.float32), frame, albedo, normal, depth)
        out[f"sample_{frame}"] = orc.sample.copy()
        out[f"accumulation_{frame}"] = orc.accumulation.copy()
    out["counters"] = orc.counters
    np.savez_compressed(G / "oracle_duck_deferred.npz", **out)
    print("oracle duck deferred: surface texels", int((depth != 0).sum()), orc.stats())


if __name__ == "__main__":
    if "--deferred-only" in sys.argv:
        deferred_pin(rf.PtFormat.loads(O.duck_pt_bytes()))
    else:
        main()
        deferred_pin(rf.PtFormat.loads(O.duck_pt_bytes()))
