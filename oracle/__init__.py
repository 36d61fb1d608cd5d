"""oracle/ -- TEST INFRASTRUCTURE ONLY (CPU restatement of the reference kernels + the reference's own
CUDA kernels compiled unmodified into oracle/_ref).  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this package; the product package
`detectron.pytorch_b200` never does."""
