"""How tests/golden/readme_vectors.json was made (run in the build container,
where /root/reference is mounted; the GPU box never reads the reference).

The reference is Go and cannot be executed here, so the vectors are the
outputs the reference itself publishes: the worked alignments and the two
M-component tables of README.md, plus the input pairs of wfa-go/seqs.txt.
The JSON was assembled by transcribing those README blocks (file:line in each
entry's "source"); tables were parsed from the markdown rows verbatim.
`stale_cells` / `cigar_stale` record where README output predates the current
source (README still documents v0.2.0 usage): see DESIGN.md section 3.
"""
