"""Compact view of one bench.py JSON line (tools/show_bench.py gpurun_out/r2_bench_n2.json)."""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(f"N={d['n_gpus']} value {d['value']} GB/s ({d['ms_per_step']} ms/step) launches {d['gpu_launches']} gate {d['parity_gate']['passed']} ranks {d['parity_gate'].get('ranks_passed')}")
print("ops:", " | ".join(f"{o['op']} {o['ms']*1e3:.0f}us {o['frac_of_hbm_roof']*100:.1f}%hbm {o['frac_of_t_bound']*100:.0f}%bound" for o in d['ops']))
r = d['roofline']; print(f"roofline: frac {r['frac']} traffic {r['traffic']} gather roof {r['l2_gather_roof_measured_TBps']} TB/s frac_binding {r['frac_of_binding_roof']}")
e = d['e2e']
if e: print(f"e2e {e['value']} GB/s ({e['ms_per_step']} ms) rows {e['gather_rows_per_step']} gather GB/s {e.get('gather_GBps')} peer roof {e.get('peer_gather_roof', {}).get('GBps')}")
t = d['train']
if t:
    print(f"train ref-shaped {t['minibatches_per_s']}/s ({t['ms_per_step_device']} ms)")
    for k in ['fused_epilogue_model', 'fused_epilogue_flat_gradients', 'live_sampler', 'live_sampler_fused_epilogue']:
        v = t.get(k, {}); print(f"  {k}: {v.get('minibatches_per_s')} /s dev {v.get('ms_per_step_device')} wall {v.get('ms_per_step_wall')} {v.get('error', '')}")
print("ref cuda kernels:", d.get('reference_cuda_kernels_same_gpu'))
for k, v in (d.get('other_workloads') or {}).items():
    if 'error' in v: print(k, 'ERROR', v['error']); continue
    if 'sampling' in v:
        for tag, s in v['sampling'].items(): print(f"{k}/{tag}: rows {s['input_rows_per_minibatch']} spmm {s['spmm_fwd_bwd_GBps_all_ranks']} GB/s {s['spmm_us_per_minibatch_max_rank']} us parity {s['parity_ok']}")
        tr = v['train_gcn_live_locality']; print(f"{k}/train: {tr.get('minibatches_per_s')} /s wall {tr.get('ms_per_step_wall')} {tr.get('error','')}")
    else:
        print(k, " ".join(f"D{w['D']}:{w['fwd_bwd_GBps_all_ranks']}" for w in v['width_sweep']), 'parity', v['parity_ok'], 'host gather', v['host_gather']['GBps_per_gpu'], 'GB/s/GPU')
if d.get('cpu_baseline'): print('cpu baseline', round(d['cpu_baseline']['value'], 3), 'GB/s', d['cpu_baseline']['cores'], 'threads')
