R4="--opt fieldsplit_u_pc_amg_eig_ratio=4 --opt fieldsplit_p_PCD_Ap_pc_amg_eig_ratio=4"
for o in "$R4 --opt fieldsplit_u_pc_amg_smooth_steps=1" "$R4 --opt fieldsplit_u_pc_amg_smooth_steps=1 --opt fieldsplit_p_PCD_Ap_pc_amg_smooth_steps=1" "--opt fieldsplit_u_pc_amg_smooth_steps=1 --opt fieldsplit_p_PCD_Ap_pc_amg_smooth_steps=1" "--opt fieldsplit_u_pc_amg_eig_ratio=3 --opt fieldsplit_p_PCD_Ap_pc_amg_eig_ratio=3" "--opt fieldsplit_u_pc_amg_eig_ratio=5 --opt fieldsplit_p_PCD_Ap_pc_amg_eig_ratio=5" "--opt fieldsplit_u_pc_amg_eig_ratio=4"; do
  timeout 200 python bench.py --steps 3 --no-cpu-baseline --no-clocks $o 2>&1 | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('$o', '|', d['fgmres_iterations'], 'its', round(d['ms_per_step'],1), 'ms', 'apply', round(d['pc_apply_only']['ms_per_apply'],3))
except Exception as e: print('$o FAILED', e)"
done
