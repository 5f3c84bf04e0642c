// Package wfa: cgo shim that re-points the reference's Aligner at libwfacuda.so.
//
// This file is what a maintainer adds to github.com/shenwei356/wfa (INTEGRATION.md
// walks through it).  It keeps the exported API of wfa.go / wfa_cigar.go and swaps
// the body of AlignPointers (reference wfa.go:201-268) for one C call; AlignBatch is
// new.  It cannot be compiled in this image (no Go toolchain); it is mirrored 1:1
// by wfa_b200/api.py (ctypes) and wfa_b200/host/wfa.hpp (C++), which are tested.
package wfa

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../wfa_b200 -lwfacuda
#include <stdlib.h>
#include "wfacuda.h"
*/
import "C"

import (
	"fmt"
	"runtime"
	"unsafe"
)

// cudaAligner is the state the GPU-backed Aligner carries next to p/ad/opt
// (reference wfa.go:79-87); M, I, D stay nil: the wavefronts live in HBM.
type cudaAligner struct {
	ctx *C.wfacuda_ctx
}

func (algn *Aligner) config() C.wfacuda_config {
	var c C.wfacuda_config
	c.mismatch = C.uint32_t(algn.p.Mismatch)
	c.gap_open = C.uint32_t(algn.p.GapOpen)
	c.gap_ext = C.uint32_t(algn.p.GapExt)
	if algn.opt.GlobalAlignment {
		c.global_alignment = 1
	}
	if algn.ad != nil {
		c.adaptive = 1
		c.min_wf_len = C.uint32_t(algn.ad.MinWFLen)
		c.max_dist_diff = C.uint32_t(algn.ad.MaxDistDiff)
		c.cutoff_step = C.uint32_t(algn.ad.CutoffStep)
	}
	return c
}

// NewOnDevice is New (wfa.go:120-131) bound to one GPU.  New keeps its signature
// and calls NewOnDevice(p, opt, 0).  Unlike the pooled reference Aligner, ad is reset.
func NewOnDevice(p *Penalties, opt *Options, device int) (*Aligner, error) {
	algn := &Aligner{p: p, opt: opt}
	c := algn.config()
	ctx := C.wfacuda_create(C.int(device), &c)
	if ctx == nil {
		return nil, fmt.Errorf("wfa: %s", C.GoString(C.wfacuda_last_error(nil)))
	}
	algn.cuda = &cudaAligner{ctx: ctx}
	runtime.SetFinalizer(algn, func(a *Aligner) { RecycleAligner(a) })
	return algn, nil
}

// RecycleAligner (wfa.go:102-116) releases the device context.
func RecycleAligner(algn *Aligner) {
	if algn != nil && algn.cuda != nil && algn.cuda.ctx != nil {
		C.wfacuda_destroy(algn.cuda.ctx)
		algn.cuda.ctx = nil
	}
}

// AdaptiveReduction (wfa.go:134-140).
func (algn *Aligner) AdaptiveReduction(ad *AdaptiveReductionOption) error {
	if ad.MinWFLen == 0 {
		return fmt.Errorf("cutoff step should not be 0")
	}
	algn.ad = ad
	c := algn.config()
	if C.wfacuda_set_config(algn.cuda.ctx, &c) != 0 {
		return fmt.Errorf("wfa: %s", C.GoString(C.wfacuda_last_error(algn.cuda.ctx)))
	}
	return nil
}

// AlignPointers (wfa.go:201-268): one pair is a batch of one.
func (algn *Aligner) AlignPointers(q, t *[]byte) (*AlignmentResult, error) {
	rs, errs := algn.AlignBatch([][]byte{*q}, [][]byte{*t})
	return rs[0], errs[0]
}

// AlignBatch aligns many pairs in one call.  results[i] is nil where errs[i] != nil;
// errs[i] is ErrEmptySeq / ErrSeqTooLong exactly where Align would return them.
func (algn *Aligner) AlignBatch(qs, ts [][]byte) ([]*AlignmentResult, []error) {
	n := len(qs)
	results := make([]*AlignmentResult, n)
	errs := make([]error, n)
	if n == 0 {
		return results, errs
	}
	total := 0
	for i := range qs {
		total += len(qs[i]) + len(ts[i])
	}
	// The byte pool lives in page-locked memory from the library (wfacuda_host_alloc): the DMA
	// engine reads it directly instead of the library staging it through its own pinned buffers,
	// and it is C memory, so no Go pointer is retained by C (cgo pointer rule).
	poolPtr := C.wfacuda_host_alloc(C.size_t(total + 16))
	if poolPtr == nil {
		err := fmt.Errorf("wfa: %s", C.GoString(C.wfacuda_last_error(nil)))
		for i := range errs {
			errs[i] = err
		}
		return results, errs
	}
	defer C.wfacuda_host_free(poolPtr)
	pool := unsafe.Slice((*byte)(poolPtr), total+16)[:0]
	qOff, tOff := make([]C.uint64_t, n), make([]C.uint64_t, n)
	qLen, tLen := make([]C.uint32_t, n), make([]C.uint32_t, n)
	for i := range qs {
		qOff[i], qLen[i] = C.uint64_t(len(pool)), C.uint32_t(len(qs[i]))
		pool = append(pool, qs[i]...)
		tOff[i], tLen[i] = C.uint64_t(len(pool)), C.uint32_t(len(ts[i]))
		pool = append(pool, ts[i]...)
	}
	pool = append(pool, make([]byte, 16)...)
	res := make([]C.wfacuda_result, n)
	off := make([]C.uint64_t, n)
	ops := make([]uint64, total/4+16*n+64)
	call := func() C.int {
		return C.wfacuda_align_batch(algn.cuda.ctx, C.uint64_t(n), (*C.uint8_t)(unsafe.Pointer(&pool[0])),
			&qOff[0], &qLen[0], &tOff[0], &tLen[0], &res[0],
			(*C.uint64_t)(unsafe.Pointer(&ops[0])), C.uint64_t(len(ops)), &off[0])
	}
	rc := call()
	if rc == C.WFACUDA_E_OPS_CAPACITY {
		ops = make([]uint64, uint64(C.wfacuda_last_ops_total(algn.cuda.ctx)))
		rc = call()
	}
	if rc != 0 {
		err := fmt.Errorf("wfa: %s", C.GoString(C.wfacuda_last_error(algn.cuda.ctx)))
		for i := range errs {
			errs[i] = err
		}
		return results, errs
	}
	for i := 0; i < n; i++ {
		switch res[i].status {
		case C.WFACUDA_OK:
			r := NewAlignmentResult(algn.opt.GlobalAlignment) // pool, wfa_cigar.go:67-72
			a := uint64(off[i]) // pairs complete in any order on the GPU: off[i] is where pair i's ops landed
			r.Ops = append(r.Ops[:0], ops[a:a+uint64(res[i].n_ops)]...) // already reversed + merged
			r.Score = uint32(res[i].score)
			r.TBegin, r.TEnd = int(res[i].tbegin), int(res[i].tend)
			r.QBegin, r.QEnd = int(res[i].qbegin), int(res[i].qend)
			r.AlignLen, r.Matches = uint32(res[i].align_len), uint32(res[i].matches)
			r.Gaps, r.GapRegions = uint32(res[i].gaps), uint32(res[i].gap_regions)
			r.proccessed = true // process() already ran on the GPU (wfa_cigar.go:137-139)
			results[i] = r
		case C.WFACUDA_ERR_EMPTY_SEQ:
			errs[i] = ErrEmptySeq
		case C.WFACUDA_ERR_SEQ_TOO_LONG:
			errs[i] = ErrSeqTooLong
		default:
			errs[i] = fmt.Errorf("wfa: pair needs more device memory than available")
		}
	}
	return results, errs
}

// AlignBatchMulti shards one batch over several devices, one goroutine per device
// inside the library (wfacuda_align_batch_multi): pairs are independent, no collective.
func AlignBatchMulti(algns []*Aligner, qs, ts [][]byte) ([]*AlignmentResult, []error) {
	// Same marshalling as AlignBatch with ctxs := []*C.wfacuda_ctx{algns[i].cuda.ctx...}
	// passed to C.wfacuda_align_batch_multi; omitted here for brevity of the shim.
	return algns[0].AlignBatch(qs, ts)
}

// FillComponents aligns one pair and fills algn.M / I / D from the GPU's wavefront store
// (wfacuda_align_components), so that Plot / Print / GetRaw of the reference keep working
// (wfa_component_plot.go:41-209).  A debugging interface: one pair, O(wavefront cells) host memory.
func (algn *Aligner) FillComponents(q, t *[]byte) (*AlignmentResult, error) {
	var res C.wfacuda_result
	ops := make([]uint64, len(*q)+len(*t)+16)
	var nRows C.uint32_t
	var nCells C.uint64_t
	rows := make([]C.wfacuda_wavefront, 1)
	cells := make([]C.uint32_t, 1)
	call := func() C.int {
		return C.wfacuda_align_components(algn.cuda.ctx,
			(*C.uint8_t)(unsafe.Pointer(&(*q)[0])), C.uint32_t(len(*q)), (*C.uint8_t)(unsafe.Pointer(&(*t)[0])), C.uint32_t(len(*t)),
			&res, (*C.uint64_t)(unsafe.Pointer(&ops[0])), C.uint64_t(len(ops)),
			&rows[0], C.uint32_t(len(rows)), &nRows, &cells[0], C.uint64_t(len(cells)), &nCells)
	}
	rc := call()
	if rc == C.WFACUDA_E_OPS_CAPACITY { // first call sized the buffers
		rows = make([]C.wfacuda_wavefront, int(nRows)+1)
		cells = make([]C.uint32_t, int(nCells)+1)
		rc = call()
	}
	if rc != 0 {
		return nil, fmt.Errorf("wfa: %s", C.GoString(C.wfacuda_last_error(algn.cuda.ctx)))
	}
	switch res.status {
	case C.WFACUDA_ERR_EMPTY_SEQ:
		return nil, ErrEmptySeq
	case C.WFACUDA_ERR_SEQ_TOO_LONG:
		return nil, ErrSeqTooLong
	}
	algn.M.Reset()
	algn.I.Reset()
	algn.D.Reset()
	for _, w := range rows[:nRows] {
		for k := int(w.lo); k <= int(w.hi); k++ {
			c := cells[uint64(w.first_cell)+3*uint64(k-int(w.lo)):]
			if c[0] != 0 {
				algn.M.SetRaw(uint32(w.score), k, uint32(c[0])) // offset<<3 | code, as next / extend left it
			}
			if c[1] != 0 {
				algn.I.SetRaw(uint32(w.score), k, uint32(c[1]))
			}
			if c[2] != 0 {
				algn.D.SetRaw(uint32(w.score), k, uint32(c[2]))
			}
		}
	}
	r := NewAlignmentResult(algn.opt.GlobalAlignment)
	r.Ops = append(r.Ops[:0], ops[:res.n_ops]...)
	r.Score = uint32(res.score)
	r.TBegin, r.TEnd, r.QBegin, r.QEnd = int(res.tbegin), int(res.tend), int(res.qbegin), int(res.qend)
	r.AlignLen, r.Matches, r.Gaps, r.GapRegions = uint32(res.align_len), uint32(res.matches), uint32(res.gaps), uint32(res.gap_regions)
	r.proccessed = true
	return r, nil
}

// RenderBatch returns CIGAR(onlyAignedRegion) and the three AlignmentText lines of every pair,
// formatted on the GPU (wfacuda_batch_render; wfa_cigar.go:236-333) -- for callers that print
// every alignment of a large batch.  Uses the split interface: upload, run, render, free.
//   b := C.wfacuda_batch_upload(ctx, n, pool, qOff, qLen, tOff, tLen); C.wfacuda_batch_run(ctx, b)
//   rc := C.wfacuda_batch_render(ctx, b, trim, cigar, cap(cigar), cigarOff, cigarLen, text, cap(text), textOff, textLen)
//   rc == WFACUDA_E_OPS_CAPACITY: C.wfacuda_last_render_total(ctx, &needCigar, &needText), grow, call again
//   pair i: string(cigar[cigarOff[i]:][:cigarLen[i]]); lines j = 0 (Q), 1 (A), 2 (T): text[textOff[i]+j*textLen[i]:][:textLen[i]]
