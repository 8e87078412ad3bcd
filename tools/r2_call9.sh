#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python tools/bnb_bench.py --instances 100 --runs lockstep:0,async:0,async:6,lockstep:6,async:14 > gpurun_out/r2c9_bnb.jsonl 2> gpurun_out/r2c9_bnb.err; cut -c1-900 gpurun_out/r2c9_bnb.jsonl; tail -5 gpurun_out/r2c9_bnb.err
