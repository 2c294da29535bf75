"""one line per bench JSON file whose path starts with the given prefix"""
import glob, json, sys
for f in sorted(glob.glob(sys.argv[1] + "*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f"{f[len(sys.argv[1]):]:60s} {d['value']:.3e} env-steps/s  {d['ms_per_step']:.2f} ms  e2e {d['e2e']['value']:.3e}  launches {d['gpu_launches']}  success {d['episode_stats']['success_rate']}")
    except Exception as exc:
        print(f, "unreadable:", exc)
