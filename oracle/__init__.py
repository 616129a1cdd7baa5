"""oracle/ — TEST INFRASTRUCTURE ONLY.

CPU restatement of (a) the reference's boundary logic for the Task -> LLM step and (b) the
numerics the reference outsources to a hosted model.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this package; the product
(agentcontrolplane_b200/, libacp_infer.so) never does.
"""
