import json,sys
d=json.loads(sys.stdin.read()); pk=d["per_kernel_ms"]
print(d["config"]["pipeline"], round(d["value"]), round(d["ms_per_step"],3), {k:v for k,v in pk.items() if k in ("hydro","ct","trace","elec_dbf","emf_z","emf_y","emf_x","update","update_ct","flux_x","flux_y","flux_z","prim_dt")})
