for bw in 32 64 128; do for ct in 32 64 128; do PG_BW_TPB=$bw PG_COMMIT_TPB=$ct python bench.py --no-cpu --no-mlp --no-nmpc --steps 20 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bw $bw commit $ct: ilqr %.4e (%.1f ms) full_batch %.1f us c4 %.3e %.3e' % (d['ilqr']['value'], d['ilqr']['ms'], d['ilqr']['full_batch']['us_per_iteration'], d['ilqr_c4']['double_parallel']['value'], d['ilqr_c4']['acrobot']['value']))"; done; done
