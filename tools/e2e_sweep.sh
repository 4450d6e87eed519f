for mib in 4 8 16 32 64; do for depth in 3 4 6; do
NTT_B200_PIPE_MIB=$mib NTT_B200_PIPE_DEPTH=$depth python bench.py --no-extras --no-cpu-baseline --steps 5 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('mib',$mib,'depth',$depth,'e2e %.0f ms %.3f sep %.0f'%(d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['separate_calls']['value']))"
done; done
