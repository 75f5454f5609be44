"""Copy the judged artefacts from gpurun_out/ into profiles/ and regenerate profiles/r1_summary.md."""
import collections, csv, json, os, shutil, subprocess, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(R, "gpurun_out"), os.path.join(R, "profiles")
def cp(src, dst):
    if os.path.exists(os.path.join(G, src)):
        shutil.copyfile(os.path.join(G, src), os.path.join(P, dst))
cp("bench_launches.csv", "r1_bench_n1_launches.csv"); cp("bench_n1.json", "r1_bench_n1.json"); cp("bench_n2.json", "r1_bench_n2.json")
cp("configs.json", "r1_configs_full_shape.json")
out = []
def launch_summary(f, title):
    f = os.path.join(P, f)
    if not os.path.exists(f): return
    rows = [r for r in csv.reader(open(f)) if len(r) > 10]
    hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try: v = float(r[vi].replace(',', ''))
        except ValueError: continue
        k = r[ki].split('(')[0][:80]
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    out.append('## %s\n(ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised: compare SHARES)\n' % title)
    out.append('total %.3f ms over %d launches\n' % (tot / 1e6, sum(a[0] for a in agg.values())))
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
        out.append('    %-82s n=%3d  %10.3f ms  %5.1f%%' % (k, c, t / 1e6, 100 * t / tot))
    out.append('')
def raw(rep, want, title):
    rep = os.path.join(G, rep)
    if not os.path.exists(rep): return
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    if len(rows) < 3: return
    hdr, units, vals = rows[0], rows[1], rows[2]
    out.append('## %s\n(ncu --set full --clock-control none; kernel: %s)\n' % (title, dict(zip(hdr, vals)).get('Kernel Name', '?')[:100]))
    for h, u, v in zip(hdr, units, vals):
        if h in want: out.append('    %-72s %-16s %s' % (h, u, v))
    out.append('')
def stalls(rep, title):
    rep = os.path.join(G, rep)
    if not os.path.exists(rep): return
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
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
launch_summary('r1_bench_n1_launches.csv', 'bench.py --steps 2 --warmup 3 (SVD config 2): every launch')
launch_summary('r1_sim_cosine_launches.csv', 'tools/profile_sim.py cosine, 8192 items x 32768 users, 4M half-star ratings, 3 builds')
launch_summary('r1_sim_pearson_baseline_launches.csv', 'tools/profile_sim.py pearson_baseline, same shape, 3 builds')
W = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread', 'launch__grid_size',
     'launch__block_size', 'launch__cluster_size', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'lts__t_sectors.sum',
     'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
     'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
     'l1tex__m_xbar2l1tex_read_bytes.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum.per_second',
     'TPC.TriageCompute.sm__pipe_tensor_subpipe_imma_cycles_active_realtime.avg', 'sm__cycles_active.avg',
     'lts__throughput.avg.pct_of_peak_sustained_elapsed']
raw('svd_prof.ncu-rep', W, 'dsgd_svd_kernel, 20 epochs of the bench workload in one launch')
stalls('svd_prof.ncu-rep', 'dsgd_svd_kernel')
raw('gemm_prof.ncu-rep', W, 'gemm_u8_tc_kernel (tcgen05 kind::i8), cosine batch: 4 accumulators, 3+3 panels, 2080 tiles x 512 k-blocks')
stalls('gemm_prof.ncu-rep', 'gemm_u8_tc_kernel')
W2 = W + ['l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
          'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct',
          'lts__t_sector_hit_rate.pct', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active']
raw('nmf_prof.ncu-rep', W2, 'nmf_pass_fused_kernel<16> (user pass), Netflix shape x0.6: 288000 users x 10620 items, 36M ratings, f=15')
stalls('nmf_prof.ncu-rep', 'nmf_pass_fused_kernel')
raw('knn_prof.ncu-rep', W2, 'knn_predict_kernel (capture of the register-select version; the committed kernel reads heads from shared memory), 2M pairs, k=40, ml-1M shape')
stalls('knn_prof.ncu-rep', 'knn_predict_kernel')
for src, dst in (('scale_n1.json', 'r1_scale_n1.json'), ('scale_n2.json', 'r1_scale_n2.json'), ('predict.json', 'r1_predict.json'),
                 ('bench_ref.json', 'r1_bench_reference_arm.json')):
    cp(src, dst)
# DRAM traffic per launch of the dominant kernels, read by bench.py (roofline.traffic)
def dram(rep):
    rep = os.path.join(G, rep)
    if not os.path.exists(rep): return None
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    tot = 0.0
    for name in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
        i = hdr.index(name)
        mul = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[units[i]]
        tot += float(vals[i].replace(',', '')) * mul
    return int(tot)
traffic = {'source': 'ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum of one launch (tools/profile_svd.py, '
                     'tools/profile_sim.py, tools/profile_nmf.py scale=0.6); see profiles/r1_summary.md'}
for key, rep in (('dsgd_svd_kernel_dram_bytes_per_launch', 'svd_prof.ncu-rep'), ('gemm_u8_tc_kernel_dram_bytes_per_launch', 'gemm_prof.ncu-rep'),
                 ('nmf_pass_fused_kernel_dram_bytes_per_launch_36M_ratings', 'nmf_prof.ncu-rep')):
    v = dram(rep)
    if v is not None: traffic[key] = v
json.dump(traffic, open(os.path.join(P, 'traffic.json'), 'w'), indent=1)
notes = open(os.path.join(P, 'r1_notes.md')).read() if os.path.exists(os.path.join(P, 'r1_notes.md')) else ''
out.append(notes)
def headline():
    """One table of every measured kernel / config, read from the JSON artefacts next to this file."""
    def load(name):
        f = os.path.join(P, name)
        return json.load(open(f)) if os.path.exists(f) else {}
    b, cfg, pred = load('r1_bench_n1.json'), load('r1_configs_full_shape.json'), load('r1_predict.json')
    s1, s2, s8 = load('r1_scale_n1.json'), load('r1_scale_n2.json'), load('r1_scale_n8.json')
    rows = ['## Headline numbers (one B200 unless stated)\n',
            '| What | Measured | Bound / fraction | Source |', '|---|---|---|---|']
    if b:
        r = b['roofline']
        rows.append('| SVD f=100, 20 epochs, ml-1M shape (bench.py) | %.3g rating-updates/s resident, %.3g end to end; kernel %.2f ms | '
                    'on-chip latency; %.2f of measured HBM peak on algorithmic bytes, DRAM traffic %.1f MB per fit | r1_bench_n1.json |'
                    % (b['value'], b['e2e']['value'], r['kernel_ms_per_launch'], r['frac'], (r['traffic'] or 0) / 1e6))
        rows.append('| reference Cython SVD.sgd on the same box | %.3g rating-updates/s (1 core) | - | r1_bench_n1.json cpu_baseline |' % b['cpu_baseline']['value'])
    c3, c4, c5 = cfg.get('c3', {}), cfg.get('c4', {}), cfg.get('c5', {})
    if c3:
        rows.append('| pearson_baseline item-item build, ml-20M shape | %.3f s (cosine %.3f s); max abs err %.2g on 20k sampled pairs | tensor (int8): see note in the file | r1_configs_full_shape.json |'
                    % (c3['sim_build_s'], c3['cosine_build_s'], c3['max_abs_err_vs_oracle_20000_sampled_pairs']))
    for name, sc in (('2', s2), ('8', s8)):
        if sc.get('c3_pearson_baseline_build_s'):
            rows.append('| same, %s GPUs (symmetric shards + NCCL exchange) | %.3f s (cosine %.3f s) | - | r1_scale_n%s.json |'
                        % (name, sc['c3_pearson_baseline_build_s'], sc['c3_cosine_build_s'], name))
    if c4:
        rows.append('| SVD++ f=20, 20 epochs, ml-10M shape | %.3f s, held-out RMSE %.5f (sequential oracle %.5f) | - | r1_configs_full_shape.json |'
                    % (c4['fit_s'], c4['heldout_rmse'], c4['oracle_heldout_rmse']))
    if c5:
        rows.append('| NMF f=15, 50 epochs, Netflix shape | %.3f s = %.3g visits/s, bit-exact vs oracle: %s | L1TEX data pipe; %.2f of measured HBM peak on algorithmic bytes | r1_configs_full_shape.json |'
                    % (c5['fit_s'], c5['rating_visits_per_s'], c5['two_epochs_bit_exact_vs_oracle'], c5['frac_of_measured_hbm']))
    if s2.get('c5_nmf_epochs_s'):
        rows.append('| same, 2 GPUs (sharded accumulators, all-gather per epoch) | %.3f s of epochs | - | r1_scale_n2.json |' % s2['c5_nmf_epochs_s'])
    if pred:
        rows.append('| k-NN estimate (k=40), ml-1M shape | %.3g pairs/s | issue (ordered selection); %.2f of HBM peak | r1_predict.json |'
                    % (pred['knn_basic']['pairs_per_s'], pred['knn_basic']['frac_of_measured_hbm']))
        rows.append('| factor-model estimate f=100 | %.3g pairs/s | L2 gather; %.2f of HBM peak | r1_predict.json |'
                    % (pred['mf_predict_f100']['pairs_per_s'], pred['mf_predict_f100']['frac_of_measured_hbm']))
    rows.append('')
    return rows
out = headline() + out
open(os.path.join(P, 'r1_summary.md'), 'w').write('# Round 1 profile summaries (B200, sm_100a)\n\nRaw artefacts next to this file: r1_*_launches.csv (ncu launch lists), r1_bench_n1.json / r1_bench_n2.json (bench lines of\nthe same build), r1_configs_full_shape.json (BASELINE.json configs 3-5 at full shape on one GPU), r1_dsgd_*.log\n(in-kernel phase counters, sb2_svd_plan_profile).\n\n' + '\n'.join(out))
print('\n'.join(out)[:3000])
