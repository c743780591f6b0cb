#!/usr/bin/env python3
"""Collects every tools/quick.py measurement of the round (gpurun_out/r02*/q_*.json) into profiles/r02_variants.md."""
import glob
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CALLS = {
    'r02': ('call 1', 'round-1 kernels as they were: baseline, the round-1 cooperative variant (tail only), 80/96-register builds, direct-mapped single-simplex tables (sm64/sm256)'),
    'r02b': ('call 2', 'first warp-synchronous path (parking at every big leaf, warp-uniform shading) + open-addressing mailbox: w80/w96 = register budgets, lm32/lm96 = leaf threshold, oh8 = overhead term, nocoop = never park'),
    'r02c': ('call 3', 'INVALID for bounce passes: the abort flag (mapped host memory) was polled on the first fetch of every warp; kept for the record'),
    'r02d': ('call 4', 'p0 = every lane for itself, w1 = warp nearest-hit traversal + per-lane shading, w2 = + warp-uniform shading, w1r = w1 at 80 registers; noaf = no adaptive fetch, nohf = no heavy-first order, fABC = rays per fetch from cost rings 0/1/2; *s = fetch-duration statistics build'),
    'r02e': ('call 5', 'wXmY: X = NTR_WARP (0 per lane, 1 warp form), Y = NTR_EXACT_MAILBOX; def = defaults of that commit (warp form for every pass of big-leaf scenes)'),
    'r02g': ('call 7', 'word-grouped item tables + sparse scan of big leaves (reverted): w0/w1/w2 = NTR_WARP'),
    'r02h': ('call 8', 'shared-memory axis tables from 9 dimensions on (noaxis = without)'),
    'r02i': ('call 9', 'exact mailbox table for opaque scenes as well (reverted; nomb = without)'),
    'r02j': ('call 10', 'quarter-size mailbox table (keys >> 2): the shipped code'),
    'r02k': ('call 11', 'thresholds of the cooperative leaves on the shipped code: md = NTR_COOP_MIN_DONE, lm = NTR_COOP_LEAF_MIN, cc = NTR_COOP_CHUNK_COST'),
    'r02l': ('call 12', 'one launch for every bounce depth, first version (mN = NTR_MERGE_FROM=N; profiles/r02_merged_queue.md)'),
    'r02m': ('call 13', 'merged bounce launch, second version'),
    'r02n': ('call 14', 'merged bounce launch, third version'),
    'r02o': ('call 15', 'merged bounce launch, fourth version (nap6/nap100 = longest pause of an idle warp 6.4 / 102 us, idle8 = 8 idle warps stay); removed afterwards'),
    'r02q': ('call 17', 'primary pass split in two launches: tiles costing more than hN mean tiles traced one 8x1 pixel row per warp by the warp-synchronous kernel, the rest beside it on a second stream; sched = NTR_TILE_SCHED=1 on a whole frame (h0_sched = the cost-sorted tile order alone).  Not kept'),
    'r02r': ('call 18', 'where the heaviest 8x4 blocks of a 1/8 share spend their time: noshadow = shadows off, depth0 = no bounces, *_stats = per-fetch durations (NTR_FETCH_STATS).  The longest block (8-9 M cycles = the whole primary pass of the share) is nearest-hit traversal alone: shadows off changes it by 2 %'),
    'r02s': ('call 19', 'software prefetch (prefetch.global.L1) of the batch record pfN items ahead in leaves of 32 items or more: +8..10 % everywhere, the longest block gets longer too -- the tail is bound by instructions, not by fetch latency.  Not kept'),
    'r02t': ('call 20', 'split primary pass with finer units: hN = cost factor, uM = pixels of a block per warp in the heavy tiles.  Even one ray per warp with 31 lanes helping does not shorten the primary pass of the share.  Not kept'),
    'r02u': ('call 21', 'settled code (def) against a single copy of the exact division + edge part of the batch test (el = NTR_EDGE_LOOP=1: 6 % fewer SASS instructions, +10 % time on config 2: the dynamic lane select costs more than the copies).  Not kept'),
    'r02v': ('call 22', 'chN = N chains of passes on ONE GPU (ntr_group_create with the device listed N times: each chain renders its interleaved tile rows with queues and launches of its own, so one chain\'s tail runs beside another\'s bulk): config 4 42.3 -> 40.4 ms with 3-4 chains, everything else flat or worse'),
    'r02w': ('call 23', 'primary pass of opaque scenes as a tracing launch (nearest hits into a buffer) + a shading launch of the same kernel (split = NTR_SPLIT_SHADE=1; frames identical, 5 GPU tests): no gain on config 2 (0.786 vs 0.779 ms; the stage switch itself cost the fused path 4 %), worse on shares and on the opaque star polytope.  Not kept -- halving the instructions each launch executes does not buy back the second launch, the hit buffer and the second tail'),
    'r02y': ('call 25', 'register budgets of the whole 3..5-D family on the settled code: cN = NTR_MIN_CTAS=N (8 -> 64 registers, 7 -> 72, 6 -> 80, 5 -> 96): whole frames want occupancy, shares and late bounce passes want registers; rKof8 = the share of rank K of 8 (13.9 .. 17.3 ms: what the 8-GPU frame waits for is its slowest rank)'),
    'r02z': ('call 26', 'both builds in one library (NTR_F_WIDE), the wide one for passes below wbN rays / pixels: no ray count separates the passes that gain from the ones that lose (0.64 M first bounces of a half frame lose, 0.62 M fourth bounces of a whole frame gain)'),
    'r02za': ('call 27', 'the build picked per pass by MEASUREMENT (wauto: frame 2 of a view ordinary, frame 3 wide, then the faster one per pass) against never (w0) and always (w1): config 4 whole 42.4 -> 42.1, 1/4 share 20.9 -> 19.9, 1/8 share 15.2 -> 14.5 ms -- but see call 28'),
    'r02zb': ('call 28', 'new = mailbox queries with the table geometry from the scene constants and column / generation by value, base = the descriptor behind a pointer (both with the wide build for shares of 4 or more GPUs only: the per-pass tuner of call 27 picked the wide build for the first bounce pass of a whole frame in the first run of this call, 44.2 ms, and was dropped)'),
    'r02zc': ('call 29', 'leaf_general without the executed re-test of the first opaque hit (new) against with it (base).  Kept'),
    'r02zd': ('call 30', 'tag mailbox for opaque single-simplex scenes, per-thread table in device memory (tN slots, pN software-pipelined, d20 = tree depth 20): tests halved, config 5 30 % slower.  Not kept'),
    'r02ze': ('call 31', 'tag mailbox, per-warp table in dynamic shared memory (tN entries), software prefetch of single-simplex records (pfK): thread-level tests -40 %, config 5 2 % slower, prefetch 5 % slower.  Not kept'),
    'r02zf': ('call 32', 'single-simplex test above 8 dimensions with aligned float4 loads and the next edge requested ahead (new) against scalar loads behind the early-exit branches (base): config 5 -1.8 %, config 5 reduced -1.1 %.  Kept'),
}


def main():
    out = ['# Round 2: every A/B measurement (`tools/quick.py`, device ms per frame, median of 5 after 3 warm-up frames, L2 flushed)\n',
           '`_w8` / `_w4` / `_w2` = one GPU rendering only the tile rows rank 0 of 8 / 4 / 2 would own (what bounds the N-GPU frame).',
           'Configs: c2 = {5,3,3} 1080p shadows; c3 = 6-D solids 1080p; c4 = {5/2,3,3} 4K reflections + transparency; c4o = its opaque',
           'variant; c4b = {5/2,5,3} 4K; c5s = 10-D soup, 16 k simplexes, 4K; c5 = 1 M simplexes.  Round-1 values: c2 0.745, c3 0.536,',
           'c4 57.6, c4o 31.1, c5s 23.3, c5 1320 ms; c4_w8 25.5 ms.\n']
    for d, (call, what) in CALLS.items():
        files = sorted(glob.glob(os.path.join(ROOT, 'gpurun_out', d, 'q_*.json')))
        if not files:
            continue
        out.append('## %s (`gpurun_out/%s/`)\n\n%s\n' % (call, d, what))
        out.append('| run | ms (median) | ms (min) |\n|---|---|---|')
        for f in files:
            try:
                j = json.load(open(f))
            except Exception:
                continue
            name = os.path.basename(f)[2:-5]
            out.append('| %s | %.3f | %.3f |' % (name, j['ms_median'], j['ms_min']))
        passes = []
        for f in sorted(glob.glob(os.path.join(ROOT, 'gpurun_out', d, 'q_c4*.err'))):
            lines = [l for l in open(f, errors='replace').read().splitlines() if l.startswith('ntr pass ms')]
            if lines:
                passes.append('`%s`: %s' % (os.path.basename(f)[2:-4], lines[-1][len('ntr pass ms:'):].strip()))
        if passes:
            out.append('\nper-pass ms (primary, bounces 1-4) and rays per bounce pass, config 4:\n')
            out += ['* ' + p for p in passes]
        out.append('')
    open(os.path.join(ROOT, 'profiles', 'r02_variants.md'), 'w').write('\n'.join(out) + '\n')


if __name__ == '__main__':
    main()
