"""TEST INFRASTRUCTURE ONLY: CPU checkers (restatement oracle + compiled reference).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package. The product package (kmer_index_b200) never does.
"""
