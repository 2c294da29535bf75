for f in gpurun_out/$1*.json; do echo -n "$f "; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('%.3e'%d['value'], '%.2f'%d['ms_per_step'], 'e2e %.3e'%d['e2e']['value'], d['gpu_launches'], d['episode_stats']['success_rate'])" 2>&1 | tail -1; done
