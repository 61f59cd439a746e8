"""TEST INFRASTRUCTURE ONLY -- CPU restatements of the reference's ensemble-dynamics path (cmrl_oracle.py: numpy fp32;
ref_port_torch.py: the reference's own torch-CPU execution, used as the timed CPU baseline; refshim.py: imports the
unmodified reference where it is mounted).  Parity is PINNED: every function is checked against golden vectors generated
by the real reference (tests/golden/) and, in the build container, against the live reference.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package; the
product (causal-mbrl_b200/) never does and fails loudly without its CUDA library (tests/test_host_logic.py checks both).
"""
