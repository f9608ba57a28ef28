import sys; sys.path.insert(0,'.')
import scrooge_b200
al = scrooge_b200.Aligner(W=64, n_gpus=1)
r = al.align_pairs(["AAAACCCCGGGGTTTT","ACGTACGT"*20], ["CCCCGGGGTTTTAAAA","ACGTACGT"*15])
print(r.edit_distances, r.cigars())
