// C++ twin of the reference's own test (wfa_test.go:30-185) and README usage
// (README.md:153-216), written against wfa_b200/host/wfa.hpp.  Unlike the
// reference's print-only test it asserts the README golden outputs.
#include <cstdio>
#include <cstdlib>
#include <string>

#include "../../wfa_b200/host/wfa.hpp"

#define CHECK(c) do { if (!(c)) { std::fprintf(stderr, "CHECK failed: %s (%s:%d)\n", #c, __FILE__, __LINE__); std::exit(1); } } while (0)

int main()
{
    wfa::Penalties p = {4, 6, 2};
    wfa::Options opt = {true};
    wfa::Aligner *algn = wfa::New(&p, &opt);
    if (!algn->ok()) { std::fprintf(stderr, "New failed: %s\n", algn->error().c_str()); return 2; }
    wfa::AdaptiveReductionOption ad = {10, 50, 1};
    CHECK(algn->AdaptiveReduction(&ad) == nullptr);
    wfa::AdaptiveReductionOption bad = {0, 50, 1};
    CHECK(algn->AdaptiveReduction(&bad) != nullptr);

    // from https://aacbb-workshop.github.io/slides/2022/WFA.ISCA.v6.pdf page15 (wfa_test.go:52-54)
    std::string q = "ACCATACTCG", t = "AGGATGCTCG";
    wfa::AlignmentResult *r = nullptr;
    CHECK(algn->Align(q, t, &r) == nullptr && r != nullptr);
    CHECK(r->CIGAR(false) == "1M2X2M1X4M");
    CHECK(r->Score == 12 && r->QBegin == 1 && r->QEnd == 10 && r->TBegin == 1 && r->TEnd == 10);
    CHECK(r->AlignLen == 10 && r->Matches == 7 && r->Gaps == 0 && r->GapRegions == 0);
    std::string Q, A, T;
    r->AlignmentText(q, t, false, &Q, &A, &T);
    CHECK(Q == "ACCATACTCG" && A == "|  || ||||" && T == "AGGATGCTCG");       // README.md:116-119
    wfa::RecycleAlignmentResult(r);

    // WFA2-lib README pair (wfa_test.go:79-85)
    CHECK(algn->Align("AGCTAGTGTCAATGGCTACTTTTCAGGTCCT", "AACTAAGTGTCGGTGGCTACTATATATCAGGTCCT", &r) == nullptr);
    CHECK(r->CIGAR(false) == "1M1X3M1I5M2X8M3I1M1X9M" && r->Score == 36);
    wfa::RecycleAlignmentResult(r);

    // errors keep their identity (wfa.go:187-193, :204-209)
    CHECK(algn->Align("", "A", &r) == wfa::ErrEmptySeq && r == nullptr);

    // batched entry point
    std::vector<wfa::AlignmentResult *> rs; std::vector<wfa::Error> es;
    CHECK(algn->AlignBatch({"C", "CG", ""}, {"C", "C", "A"}, &rs, &es) == nullptr);
    CHECK(es[0] == nullptr && rs[0]->CIGAR(false) == "1M" && es[2] == wfa::ErrEmptySeq && rs[2] == nullptr && es[1] == nullptr);
    for (auto *x : rs) wfa::RecycleAlignmentResult(x);
    wfa::RecycleAligner(algn);

    // semi-global with trimming (README.md:18-27, wfa-go -g -t)
    wfa::Options semi = {false};
    algn = wfa::New(&p, &semi);
    q = "Bioinformatics helps Biology"; t = "We learn bioinformatics to help biologists";
    CHECK(algn->Align(q, t, &r) == nullptr);
    CHECK(r->CIGAR(false) == "9I1X14M3I4M1D1M1X5M1X3I" && r->CIGAR(true) == "14M3I4M1D1M1X5M");
    CHECK(r->Score == 32 && r->QBegin == 2 && r->QEnd == 27 && r->TBegin == 11 && r->TEnd == 38);
    r->AlignmentText(q, t, false, &Q, &A, &T);
    CHECK(Q == "---------Bioinformatics ---helps Biology---");
    CHECK(T == "We learn bioinformatics to help- biologists");
    wfa::RecycleAlignmentResult(r);

    // the same strings formatted on the GPU for a whole batch (wfa-go -g -t prints these)
    std::vector<std::string> cg, bQ, bA, bT; std::vector<wfa::Error> bes;
    CHECK(algn->AlignBatchRendered({q, "", "ACGT"}, {t, "A", "ACGT"}, true, &cg, &bQ, &bA, &bT, &bes) == nullptr);
    CHECK(cg[0] == "14M3I4M1D1M1X5M" && bQ[0] == "ioinformatics ---helps Biolog" && bT[0] == "ioinformatics to help- biolog");
    CHECK(bA[0].size() == bQ[0].size() && bes[1] == wfa::ErrEmptySeq && cg[1].empty() && cg[2] == "4M" && bA[2] == "||||");
    wfa::RecycleAligner(algn);

    // Aligner.M / I / D from the GPU's wavefront store: the known-answer trace of README.md:101-124
    algn = wfa::New(&p, &opt);
    wfa::Aligner::Components comps;
    CHECK(algn->AlignComponents("ACCATACTCG", "AGGATGCTCG", &r, &comps) == nullptr && r->CIGAR(false) == "1M2X2M1X4M");
    CHECK(comps.GetRaw(0, 0, 0) == (1u << 3 | 6u) && comps.GetRaw(0, 8, 0) == (5u << 3 | 5u));     // M[0][0] = Match 1, M[8][0] = Mismatch 5 (after extend)
    CHECK(comps.GetRaw(2, 8, -1) == (1u << 3 | 3u) && comps.GetRaw(1, 8, 1) == (2u << 3 | 1u));    // D[8][-1] = DelOpen 1, I[8][1] = InsOpen 2
    CHECK(comps.GetRaw(0, 12, 0) == (10u << 3 | 5u) && comps.GetRaw(0, 6, 0) == 0u);
    wfa::RecycleAlignmentResult(r);
    wfa::RecycleAligner(algn);
    std::printf("host API ok\n");
    return 0;
}
