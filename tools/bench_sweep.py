"""Run bench.py under a list of environment variants and print one compact line per variant (GPU box helper).

    python tools/bench_sweep.py --reads 2000000 "HSA_B200_MINB=2" "HSA_B200_MINB=3" "HSA_B200_MINB=4 HSA_B200_ARENA_CAP=2048"
"""
import json
import os
import subprocess
import sys


def main():
    args = sys.argv[1:]
    reads = "2000000"
    if args and args[0] == "--reads":
        reads, args = args[1], args[2:]
    for variant in args or [""]:
        env = dict(os.environ)
        for kv in variant.split():
            k, v = kv.split("=", 1)
            env[k] = v
        p = subprocess.run([sys.executable, "bench.py", "--reads", reads, "--steps", "2", "--warmup", "1",
                            "--no-cpu-baseline", "--no-probe"], env=env, capture_output=True, text=True)
        try:
            j = json.loads(p.stdout.strip().splitlines()[-1])
            print(f"[{variant or 'default'}] value={j['value'] / 1e6:.2f} M/s ms={j['ms_per_step']:.1f} e2e={j['e2e']['value'] / 1e6:.2f} M/s "
                  f"achieved={j['roofline']['achieved']:.0f} GB/s strict={j['reads_needing_strict_rerun']} same={j['device_vs_host_path_identical']} "
                  f"launches={j['gpu_launches']}", flush=True)
        except Exception as e:  # noqa: BLE001
            print(f"[{variant}] FAILED rc={p.returncode} {e}\n{p.stderr[-1500:]}", flush=True)


if __name__ == "__main__":
    main()
