"""A/B builds of libmaai_ntxent.so for tuning (run HERE: nvcc cross-compiles without a GPU), and
the runner that times them on the GPU box.

    python tools/ab_variants.py build  name=DEF1,DEF2 ...   # -> multimodal-active-ai_b200/variants/name.so
    python tools/ab_variants.py run [B d iters]             # on the B200: times every variant

`run` starts one process per variant (MAAI_DEBUG_LIB selects the library) and prints CUDA-event
medians of the forward and backward C-ABI calls.  Variants built with MAAI_ABL != 0 compute wrong
results on purpose (timing-only ablations).
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VDIR = os.path.join(ROOT, "multimodal-active-ai_b200", "variants")


def build(specs):
    import importlib.util
    spec = importlib.util.spec_from_file_location("maai_build", os.path.join(ROOT, "multimodal-active-ai_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    os.makedirs(VDIR, exist_ok=True)
    procs = []
    for s in specs:
        name, _, defs = s.partition("=")
        defs = tuple(d for d in defs.split(",") if d)
        out = os.path.join(VDIR, name + ".so")
        cmd = [mod._nvcc(), *mod.NVCC_FLAGS, *[f"-D{d}" for d in defs], "-o", out,
               *[os.path.join(mod.CSRC, x) for x in mod.SOURCES]]
        procs.append((name, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for name, p in procs:
        out, _ = p.communicate()
        print(f"{name}: rc={p.returncode}")
        if p.returncode != 0:
            print(out)
            raise SystemExit(1)


def run(args):
    libs = sorted(f for f in os.listdir(VDIR) if f.endswith(".so"))
    for f in libs:
        env = dict(os.environ, MAAI_DEBUG_LIB=os.path.join(VDIR, f))
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "quick_time.py"), *args], env=env,
                           capture_output=True, text=True)
        print(r.stdout.strip() if r.returncode == 0 else f"{f}: FAILED\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}", flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build(sys.argv[2:])
    else:
        run(sys.argv[2:])
