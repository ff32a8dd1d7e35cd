import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.chdir(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lisa_b200.frontend as fe, lisa_b200.rt as rt
sc = fe.parse_scene("scenes/cornell_c2.rto")
for i in range(8):
    t0 = time.perf_counter(); R = rt.Renderer.from_scene(sc); t1 = time.perf_counter()
    R.render_subframes(i, 1, 50); t2 = time.perf_counter()
    img = R.read_accum(); t3 = time.perf_counter()
    st = R.stats()
    R.close(); t4 = time.perf_counter()
    print("create %.1f ms (upload %.2f build %.2f) | render %.1f ms (device %.1f) | read %.1f ms | destroy %.1f ms" % ((t1-t0)*1e3, st["upload_ms"], st["bvh_build_ms"], (t2-t1)*1e3, st["last_render_ms"], (t3-t2)*1e3, (t4-t3)*1e3))
