# round 2: many small graphs (two per CTA) through the tensor-core products of the Sinkhorn stage
timeout 100 python -m pytest tests/test_gpu_mgm_solver.py -q --tb=short -x --timeout 60 -k many_graphs 2>&1 | tail -4 | cut -c1-300
