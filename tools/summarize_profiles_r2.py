"""Turn the round-2 ncu reports of tools/gpu_profile_r2.sh (gpurun_out/r2_*) into profiles/r2_summary.md, copy the launch
list next to it and refresh profiles/traffic.json (read by bench.py for roofline.traffic)."""
import collections, csv, json, os, shutil, subprocess
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(R, "gpurun_out"), os.path.join(R, "profiles")
out = []


def ncu_csv(rep, page):
    txt = subprocess.run(['ncu', '-i', os.path.join(G, rep), '--page', page, '--csv'], capture_output=True, text=True).stdout
    return list(csv.reader(txt.splitlines()))


def launch_summary(f, title):
    src = os.path.join(G, f)
    if not os.path.exists(src):
        return
    shutil.copyfile(src, os.path.join(P, f))
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(',', ''))
        except ValueError:
            continue
        a = agg.setdefault(r[ki].split('(')[0][:80], [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    out.append('## %s\n(ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised: compare SHARES)\n' % title)
    out.append('total %.3f ms over %d launches\n' % (tot / 1e6, sum(a[0] for a in agg.values())))
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
        out.append('    %-82s n=%3d  %10.3f ms  %5.1f%%' % (k, c, t / 1e6, 100 * t / tot))
    out.append('')


W = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread', 'launch__grid_size',
     'launch__block_size', 'launch__cluster_size', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'lts__t_sectors.sum',
     'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
     'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
     'TPC.TriageCompute.sm__pipe_tensor_subpipe_imma_cycles_active_realtime.avg', 'sm__cycles_active.avg',
     'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
     'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
     'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum']


def raw(rep, title, row=0):
    if not os.path.exists(os.path.join(G, rep)):
        return None
    rows = ncu_csv(rep, 'raw')
    if len(rows) < 3:
        return None
    hdr, units, vals = rows[0], rows[1], rows[2 + row]
    d = dict(zip(hdr, vals))
    out.append('## %s\n(ncu --set full --clock-control none; kernel: %s)\n' % (title, d.get('Kernel Name', '?')[:110]))
    for h, u, v in zip(hdr, units, vals):
        if h in W:
            out.append('    %-76s %-16s %s' % (h, u, v))
    out.append('')
    tot = 0.0
    for name in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
        i = hdr.index(name)
        tot += float(vals[i].replace(',', '')) * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[units[i]]
    return int(tot)


def stalls(rep, title):
    if not os.path.exists(os.path.join(G, rep)):
        return
    rows = ncu_csv(rep, 'source')
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) >= len(hdr) and r[ix['# Samples']].strip().isdigit()]
    st = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    tot = sum(int(r[ix['# Samples']] or 0) for r in data)
    out.append('## %s: warp-stall samples (source page)\n' % title)
    for s, v in sorted(((s, sum(int(r[ix[s]] or 0) for r in data)) for s in st), key=lambda kv: -kv[1])[:8]:
        out.append('    %-28s %8d %5.1f%%' % (s, v, 100 * v / max(tot, 1)))
    out.append('    top instructions:')
    for r in sorted(data, key=lambda r: -int(r[ix['# Samples']] or 0))[:8]:
        out.append('      %-60s samples=%s' % (r[ix['Source']][:60], r[ix['# Samples']]))
    out.append('')


launch_summary('r2_bench_launches.csv', 'bench.py --no-secondary --no-cpu-baseline --steps 2 --warmup 3 (SVD config 2): every launch')
traffic = {'source': 'profiles/r2_summary.md: ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of '
                     'the committed kernel (tools/gpu_profile_r2.sh); not measured inside the bench run'}
t = raw('r2_svd_prof.ncu-rep', 'dsgd_svd_kernel<8, 4, 1, 1, 1, 0>: 20 epochs of the bench workload (config 2) in one launch')
if t: traffic['dsgd_svd_kernel_dram_bytes_per_launch'] = t
stalls('r2_svd_prof.ncu-rep', 'dsgd_svd_kernel')
t = raw('r2_gemm_prof.ncu-rep', 'gemm_u8_tc_kernel (tcgen05 kind::i8, cta_group::2), cosine at 8192 x 32768: one launch of 2 accumulators')
if t: traffic['gemm_u8_tc_kernel_dram_bytes_per_launch'] = t
stalls('r2_gemm_prof.ncu-rep', 'gemm_u8_tc_kernel')
t = raw('r2_simrows_prof.ncu-rep', 'sim_rows_kernel<3> (general fp64 path, pearson_baseline) at 8192 x 32768, 4M half-star ratings')
if t: traffic['sim_rows_kernel_dram_bytes_per_launch_8192x32768'] = t
stalls('r2_simrows_prof.ncu-rep', 'sim_rows_kernel')
t = raw('r2_knn_prof.ncu-rep', 'knn_predict_kernel (threshold select, the committed kernel): 500k pairs, k = 40, ml-1M shape')
if t: traffic['knn_predict_kernel_dram_bytes_per_launch_500k_pairs'] = t
stalls('r2_knn_prof.ncu-rep', 'knn_predict_kernel')
t = raw('r2_nmf_prof.ncu-rep', 'nmf_pass_fused_kernel<16> (user pass), Netflix shape x0.6: 288000 users x 10620 items, 36M ratings, f=15')
if t: traffic['nmf_pass_fused_kernel_dram_bytes_per_launch_36M_ratings'] = t
stalls('r2_nmf_prof.ncu-rep', 'nmf_pass_fused_kernel')
json.dump(traffic, open(os.path.join(P, 'traffic.json'), 'w'), indent=1)
for f in ('r2_svd_plain.log', 'r2_gemm_plain.log', 'r2_simrows_plain.log', 'r2_knn_plain.log', 'r2_nmf_plain.log'):
    if os.path.exists(os.path.join(G, f)):
        out.append('### %s (the same command without ncu)\n\n    ' + open(os.path.join(G, f)).read().strip().replace('\n', '\n    ') + '\n')
        out[-1] = out[-1] % f
open(os.path.join(P, 'r2_ncu_summary.md'), 'w').write('# Round 2 ncu summaries (B200, sm_100a; committed kernels, tools/gpu_profile_r2.sh)\n\n' + '\n'.join(out))
print('\n'.join(out))
