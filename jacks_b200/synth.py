"""Deterministic synthetic screens for parity tests and throughput runs.

Host-side numpy only.  Shapes and distributions follow SURVEY.md section 8(d):
the "Avana shape" is 18,000 genes x 4 gRNAs x 400 cell lines with a common
control, 10 % essential genes and 1,000 negative-control genes; pseudo-genes are
random guide sets drawn from the control genes the way the reference's
``createPseudoNonessGenes`` draws them (reference jacks/jacks/jacks_io.py:256-264).

Nothing here touches the GPU; the arrays are what a caller of
``jacks.infer.inferJACKS`` (reference jacks/jacks/infer.py:18) would hold.
"""
import numpy as np

AVANA_SEED = 20180731


def make_condensed_screen(n_genes=18000, guides_per_gene=4, n_lines=400,
                          seed=AVANA_SEED, frac_essential=0.10, n_ctrl_genes=1000,
                          guide_counts=None):
    """Build (gene_index, testdata, ctrldata, ctrl_genes) at the condensed level.

    testdata/ctrldata are C-contiguous float64 ``[N, L, 2]`` (mean, sd) exactly as
    ``collateTestControlSamples`` (reference preprocess.py:121-126) hands them to
    ``inferJACKS``.  The control is common to all lines, so every ctrldata column
    is identical, as in the reference's ``--common_ctrl_sample`` path.

    guide_counts: optional int array [n_genes] for ragged screens; default all
    ``guides_per_gene``.
    """
    rng = np.random.default_rng(seed)
    if guide_counts is None:
        guide_counts = np.full(n_genes, guides_per_gene, dtype=np.int64)
    guide_counts = np.asarray(guide_counts, dtype=np.int64)
    n_genes = len(guide_counts)
    offsets = np.zeros(n_genes + 1, dtype=np.int64)
    np.cumsum(guide_counts, out=offsets[1:])
    N, L = int(offsets[-1]), int(n_lines)

    essential = rng.random(n_genes) < frac_essential
    w = np.where(essential[:, None],
                 rng.normal(-1.5, 0.5, size=(n_genes, L)),
                 rng.normal(0.0, 0.15, size=(n_genes, L)))
    x = rng.normal(1.0, 0.4, size=N)
    sd_test = rng.uniform(0.1, 0.6, size=(N, L))
    sd_ctrl = rng.uniform(0.05, 0.3, size=N)
    ctrl_mean = rng.normal(0.0, 1.0, size=N)
    gene_of_row = np.repeat(np.arange(n_genes), guide_counts)
    noise = rng.normal(0.0, 1.0, size=(N, L))
    y = x[:, None] * w[gene_of_row] + noise * np.sqrt(sd_test ** 2 + sd_ctrl[:, None] ** 2)

    testdata = np.empty((N, L, 2), dtype=np.float64)
    testdata[:, :, 0] = ctrl_mean[:, None] + y
    testdata[:, :, 1] = sd_test
    ctrldata = np.empty((N, L, 2), dtype=np.float64)
    ctrldata[:, :, 0] = ctrl_mean[:, None]
    ctrldata[:, :, 1] = sd_ctrl[:, None]

    names = ['G%05d' % i for i in range(n_genes)]
    gene_index = {names[i]: list(range(int(offsets[i]), int(offsets[i + 1]))) for i in range(n_genes)}
    noness = np.flatnonzero(~essential)
    n_ctrl = min(n_ctrl_genes, len(noness))
    ctrl_ids = np.sort(rng.choice(noness, size=n_ctrl, replace=False)) if n_ctrl > 0 else np.zeros(0, dtype=np.int64)
    ctrl_genes = [names[i] for i in ctrl_ids]
    return gene_index, testdata, ctrldata, ctrl_genes


def make_pseudo_index_csr(gene_offsets, ctrl_gene_ids, n_pseudo, seed=AVANA_SEED + 1):
    """Vectorised pseudo-gene index sets in CSR form for throughput runs.

    Same distribution as the reference's createPseudoNonessGenes
    (jacks_io.py:256-264): the guide count is copied from a uniformly drawn gene,
    then each guide is drawn by first picking a control gene uniformly and then a
    guide of it uniformly, with replacement.  (Parity runs use the host port in
    ``jacks_b200.jacks_io.createPseudoNonessGenes`` which consumes Python's
    ``random`` stream exactly like the reference.)
    Returns (pseudo_offsets int64[n_pseudo+1], pseudo_rows int32[sum G]).
    """
    rng = np.random.default_rng(seed)
    gene_offsets = np.asarray(gene_offsets, dtype=np.int64)
    counts = np.diff(gene_offsets)
    n_genes = len(counts)
    ctrl_gene_ids = np.asarray(ctrl_gene_ids, dtype=np.int64)
    g_of = counts[rng.integers(0, n_genes, size=n_pseudo)]
    offs = np.zeros(n_pseudo + 1, dtype=np.int64)
    np.cumsum(g_of, out=offs[1:])
    total = int(offs[-1])
    pick_gene = ctrl_gene_ids[rng.integers(0, len(ctrl_gene_ids), size=total)]
    within = (rng.random(total) * counts[pick_gene]).astype(np.int64)
    rows = (gene_offsets[pick_gene] + within).astype(np.int32)
    return offs, rows


def make_count_screen(n_genes=2000, guides_per_gene=4, n_lines=20, n_reps=2, seed=AVANA_SEED,
                      frac_essential=0.10, ctrl_name='pDNA'):
    """Count-level synthetic screen ("count-level input").

    Returns (guide_ids, gene_ids, colnames, samples, counts[int64 N x C]) where the
    first ``n_reps`` columns are the control sample ``ctrl_name`` and the rest are
    ``n_lines`` cell lines with ``n_reps`` replicates each.
    """
    rng = np.random.default_rng(seed)
    N = n_genes * guides_per_gene
    essential = rng.random(n_genes) < frac_essential
    w = np.where(essential[:, None], rng.normal(-1.5, 0.5, size=(n_genes, n_lines)),
                 rng.normal(0.0, 0.15, size=(n_genes, n_lines)))
    x = rng.normal(1.0, 0.4, size=N)
    lam = rng.lognormal(np.log(500.0), 0.6, size=N)
    gene_of_row = np.repeat(np.arange(n_genes), guides_per_gene)
    cols, samples, blocks = [], [], []
    for r in range(n_reps):
        depth = rng.uniform(0.7, 1.3)
        blocks.append(rng.poisson(lam * depth))
        cols.append('%s_rep%d' % (ctrl_name, r + 1)); samples.append(ctrl_name)
    for l in range(n_lines):
        for r in range(n_reps):
            depth = rng.uniform(0.7, 1.3)
            lfc = x * w[gene_of_row, l] + rng.normal(0.0, 0.25, size=N)
            blocks.append(rng.poisson(lam * depth * np.exp2(lfc)))
            cols.append('LINE%03d_rep%d' % (l, r + 1)); samples.append('LINE%03d' % l)
    counts = np.stack(blocks, axis=1).astype(np.int64)
    alphabet = np.array(list('ACGT'))
    guide_ids = [''.join(alphabet[rng.integers(0, 4, size=20)]) + '_%06d' % i for i in range(N)]
    gene_ids = ['G%05d' % g for g in gene_of_row]
    return guide_ids, gene_ids, cols, samples, counts


# Sample layout of the reference's full example screen, jacks/example/example_repmap.tab (the count file itself is not in
# the reference repository): 53 replicate columns -> the control CTRL (2 replicates) and 16 cell lines.
EXAMPLE_LAYOUT = (('CTRL', 2), ('MOLM', 2), ('MV411', 2), ('HL60', 2), ('OCIAML2', 2), ('OCIAML3', 2), ('A375', 3), ('A2058', 3),
                  ('CO205', 3), ('HuPT3', 3), ('HuPT4', 3), ('MB415', 3), ('MB436', 3), ('MB453', 3), ('SW1990', 2), ('SW48', 3),
                  ('HT29', 12))


def make_example_count_screen(n_genes=18000, guides_per_gene=5, layout=EXAMPLE_LAYOUT, seed=AVANA_SEED + 7, frac_essential=0.10):
    """BASELINE config 2 stand-in: a count-level screen with the sample layout of jacks/example (control first) and
    ``n_genes * guides_per_gene`` guides (90,000 by default).  Same count model as make_count_screen.
    Returns (guide_ids, gene_ids, colnames, samples, counts[int64 N x C])."""
    rng = np.random.default_rng(seed)
    N = n_genes * guides_per_gene
    lines = [name for name, _ in layout[1:]]
    essential = rng.random(n_genes) < frac_essential
    w = np.where(essential[:, None], rng.normal(-1.5, 0.5, size=(n_genes, len(lines))),
                 rng.normal(0.0, 0.15, size=(n_genes, len(lines))))
    x = rng.normal(1.0, 0.4, size=N)
    lam = rng.lognormal(np.log(500.0), 0.6, size=N)
    gene_of_row = np.repeat(np.arange(n_genes), guides_per_gene)
    cols, samples, blocks = [], [], []
    for li, (name, reps) in enumerate(layout):
        for r in range(reps):
            depth = rng.uniform(0.7, 1.3)
            if li == 0:
                blocks.append(rng.poisson(lam * depth))
            else:
                lfc = x * w[gene_of_row, li - 1] + rng.normal(0.0, 0.25, size=N)
                blocks.append(rng.poisson(lam * depth * np.exp2(lfc)))
            cols.append('%s_rep_%d' % (name, r + 1)); samples.append(name)
    counts = np.stack(blocks, axis=1).astype(np.int64)
    alphabet = np.array(list('ACGT'))
    order = rng.permutation(N)                   # guide ids do not sort in gene order: the loader's string sort matters
    guide_ids = [''.join(alphabet[rng.integers(0, 4, size=12)]) + '_%06d' % order[i] for i in range(N)]
    gene_ids = ['G%05d' % g for g in gene_of_row]
    return guide_ids, gene_ids, cols, samples, counts


def write_count_screen(prefix, guide_ids, gene_ids, cols, samples, counts, ctrl_genes=()):
    """Write the count table, the replicate map and (optionally) a control-gene list the way run_JACKS.py reads them."""
    import pandas as pd
    df = pd.DataFrame(counts, columns=cols)
    df.insert(0, 'Gene', gene_ids)
    df.insert(0, 'sgRNA', guide_ids)
    df.to_csv(prefix + '_counts.tab', sep='\t', index=False)
    with open(prefix + '_repmap.tab', 'w') as f:
        f.write('Replicate\tSample\n')
        for c, sname in zip(cols, samples):
            f.write('%s\t%s\n' % (c, sname))
    if len(ctrl_genes):
        with open(prefix + '_ctrl_genes.txt', 'w') as f:
            for g in ctrl_genes:
                f.write('%s\n' % g)
    return prefix + '_counts.tab', prefix + '_repmap.tab', prefix + '_ctrl_genes.txt'
