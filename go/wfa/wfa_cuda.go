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
// (reference wfa.go:79-87); M, I, D stay nil until FillComponents: the wavefronts live in HBM.
// All buffers that cross the C boundary are page-locked C memory from wfacuda_host_alloc, kept and
// grown across calls (allocating page-locked memory costs more than aligning a batch): the DMA
// engines read and write them directly, and no Go pointer is retained by C (cgo pointer rule).
type cudaAligner struct {
	ctx *C.wfacuda_ctx
	buf [8]pinned // pool, qOff, qLen, tOff, tLen, results, opsOff, ops
}

type pinned struct {
	p   unsafe.Pointer
	cap int // bytes
}

// reserve returns page-locked memory of at least n bytes (contents undefined).
func (b *pinned) reserve(n int) (unsafe.Pointer, error) {
	if n > b.cap {
		if b.p != nil {
			C.wfacuda_host_free(b.p)
			b.p, b.cap = nil, 0
		}
		want := n + n/4 + 4096
		b.p = C.wfacuda_host_alloc(C.size_t(want))
		if b.p == nil {
			return nil, fmt.Errorf("wfa: %s", C.GoString(C.wfacuda_last_error(nil)))
		}
		b.cap = want
	}
	return b.p, nil
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
		for i := range algn.cuda.buf {
			if algn.cuda.buf[i].p != nil {
				C.wfacuda_host_free(algn.cuda.buf[i].p)
				algn.cuda.buf[i] = pinned{}
			}
		}
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
	return alignBatchOn([]*Aligner{algn}, qs, ts)
}

// AlignBatchMulti shards one batch over several devices, one Aligner (ctx) per device: the
// library cuts length-binned, work-balanced shards and drives every device from its own host
// thread (wfacuda_align_batch_multi); pairs are independent, there is no collective.  The
// buffers are those of algns[0].
func AlignBatchMulti(algns []*Aligner, qs, ts [][]byte) ([]*AlignmentResult, []error) {
	return alignBatchOn(algns, qs, ts)
}

func alignBatchOn(algns []*Aligner, qs, ts [][]byte) ([]*AlignmentResult, []error) {
	algn := algns[0]
	n := len(qs)
	results := make([]*AlignmentResult, n)
	errs := make([]error, n)
	fail := func(err error) ([]*AlignmentResult, []error) {
		for i := range errs {
			errs[i] = err
		}
		return results, errs
	}
	if n == 0 {
		return results, errs
	}
	total := 0
	for i := range qs {
		total += len(qs[i]) + len(ts[i])
	}
	// page-locked buffers of the Aligner, reused call after call
	b := &algn.cuda.buf
	sizes := [7]int{total + 16, 8 * n, 4 * n, 8 * n, 4 * n, int(unsafe.Sizeof(C.wfacuda_result{})) * n, 8 * n}
	var ptr [8]unsafe.Pointer
	for i, sz := range sizes {
		p, err := b[i].reserve(sz)
		if err != nil {
			return fail(err)
		}
		ptr[i] = p
	}
	opsWords := total/4 + 16*n + 64
	if b[7].cap/8 > opsWords {
		opsWords = b[7].cap / 8
	}
	p7, err := b[7].reserve(8 * opsWords)
	if err != nil {
		return fail(err)
	}
	ptr[7] = p7
	pool := unsafe.Slice((*byte)(ptr[0]), total+16)
	qOff, qLen := unsafe.Slice((*C.uint64_t)(ptr[1]), n), unsafe.Slice((*C.uint32_t)(ptr[2]), n)
	tOff, tLen := unsafe.Slice((*C.uint64_t)(ptr[3]), n), unsafe.Slice((*C.uint32_t)(ptr[4]), n)
	at := 0
	for i := range qs {
		qOff[i], qLen[i] = C.uint64_t(at), C.uint32_t(len(qs[i]))
		at += copy(pool[at:], qs[i])
		tOff[i], tLen[i] = C.uint64_t(at), C.uint32_t(len(ts[i]))
		at += copy(pool[at:], ts[i])
	}
	for i := at; i < total+16; i++ {
		pool[i] = 0
	}
	// one ctx per device; the array itself is C memory too
	ctxs := (*[64]*C.wfacuda_ctx)(C.malloc(C.size_t(len(algns)) * C.size_t(unsafe.Sizeof((*C.wfacuda_ctx)(nil)))))
	defer C.free(unsafe.Pointer(ctxs))
	for i, a := range algns {
		ctxs[i] = a.cuda.ctx
	}
	call := func() C.int {
		if len(algns) == 1 {
			return C.wfacuda_align_batch(algn.cuda.ctx, C.uint64_t(n), (*C.uint8_t)(ptr[0]),
				&qOff[0], &qLen[0], &tOff[0], &tLen[0], (*C.wfacuda_result)(ptr[5]),
				(*C.uint64_t)(ptr[7]), C.uint64_t(opsWords), (*C.uint64_t)(ptr[6]))
		}
		return C.wfacuda_align_batch_multi(&ctxs[0], C.int(len(algns)), C.uint64_t(n), (*C.uint8_t)(ptr[0]),
			&qOff[0], &qLen[0], &tOff[0], &tLen[0], (*C.wfacuda_result)(ptr[5]),
			(*C.uint64_t)(ptr[7]), C.uint64_t(opsWords), (*C.uint64_t)(ptr[6]))
	}
	rc := call()
	if rc == C.WFACUDA_E_OPS_CAPACITY {
		opsWords = int(C.wfacuda_last_ops_total(algn.cuda.ctx))
		if ptr[7], err = b[7].reserve(8 * opsWords); err != nil {
			return fail(err)
		}
		rc = call()
	}
	if rc != 0 {
		return fail(fmt.Errorf("wfa: %s", C.GoString(C.wfacuda_last_error(algn.cuda.ctx))))
	}
	res := unsafe.Slice((*C.wfacuda_result)(ptr[5]), n)
	off := unsafe.Slice((*C.uint64_t)(ptr[6]), n)
	ops := unsafe.Slice((*uint64)(ptr[7]), opsWords)
	for i := 0; i < n; i++ {
		switch res[i].status {
		case C.WFACUDA_OK:
			r := NewAlignmentResult(algn.opt.GlobalAlignment) // pool, wfa_cigar.go:67-72
			a := uint64(off[i])                               // pairs complete in any order on the GPU: off[i] is where pair i's ops landed
			r.Ops = append(r.Ops[:0], ops[a:a+uint64(res[i].n_ops)]...) // already reversed + merged; copied into Go memory
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
	// the GPU-backed Aligner has no host-side components until somebody asks for them
	if algn.M == nil {
		algn.M, algn.I, algn.D = NewComponent(), NewComponent(), NewComponent()
		algn.M.IsM = true
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
