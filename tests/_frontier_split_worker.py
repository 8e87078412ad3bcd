"""Launched by tests/test_examples.py under torch.distributed.run (2 ranks, gloo): runs examples/frontier_split.py with the
oracle-backed stand-in engine, so the example's rank / world / exchange plumbing is exercised without GPUs."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, HERE); sys.path.insert(0, os.path.join(os.path.dirname(HERE), "examples"))
import fake_engine                      # noqa: E402
from miosqp_b200 import engine          # noqa: E402

engine.BatchedQP = fake_engine.FakeBatchedQP
engine.solve_multi = fake_engine.solve_multi
import frontier_split                   # noqa: E402

frontier_split.main(sys.argv[1:])
