// Drop-in for frcfrc/unifrac.go's unifrac() in fluhus/frackyfrac (see INTEGRATION.md; adjust the #cgo paths to
// where this repo is checked out).  NOT compiled in this image: there is no Go toolchain.
package main

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../frackyfrac_b200/_build -lfrcfrc_cuda -Wl,-rpath,${SRCDIR}/../../frackyfrac_b200/_build
#include <stdlib.h>
#include "frcfrc_cuda.h"
*/
import "C"

import (
	"fmt"
	"iter"
	"os"
	"runtime"
	"unsafe"

	"github.com/fluhus/biostuff/formats/newick"
)

// flatten mirrors enumerateNodes (unifrac.go:127-133) and the treeDists loop (:117-120):
// pre-order ids, parent[] and length[] (root included), plus name -> leaf ids for A2/A6.
func flatten(tree *newick.Node) (parent []C.int32_t, length []C.double, leaves map[string][]int32) {
	ids := map[*newick.Node]int32{}
	leaves = map[string][]int32{}
	var walk func(n *newick.Node, p int32)
	walk = func(n *newick.Node, p int32) {
		id := int32(len(parent))
		ids[n] = id
		parent = append(parent, C.int32_t(p))
		length = append(length, C.double(n.Distance))
		if len(n.Children) == 0 {
			leaves[n.Name] = append(leaves[n.Name], id) // only leaves carry abundance (unifrac.go:38-43)
		}
		for _, c := range n.Children {
			walk(c, id)
		}
	}
	walk(tree, -1)
	return
}

// unifrac keeps the reference signature (unifrac.go:97) and contract: an ordered stream of
// float64 in IterPairs order; stopping early is allowed.
func unifrac(abnd []map[string]float64, tree *newick.Node, weighted bool) iter.Seq[float64] {
	fmt.Fprintln(os.Stderr, "Converting abundances")
	parent, length, leaves := flatten(tree)

	rowPtr := make([]C.int64_t, 1, len(abnd)+1)
	var col []C.int32_t
	var val []C.double
	for _, m := range abnd {
		for name, v := range m {
			if !(v > 0) { // unifrac.go:40 only takes a > 0 (the parsers never deliver anything else)
				continue
			}
			for _, leaf := range leaves[name] { // a name on several leaves feeds all of them (A6);
				col = append(col, C.int32_t(leaf)) // names of internal nodes only: dropped (A2)
				val = append(val, C.double(v))
			}
		}
		rowPtr = append(rowPtr, C.int64_t(len(col)))
	}

	// cgo: the C side copies everything before frc_create returns, so Go slices are passed directly.  The
	// structs below hold Go pointers and are themselves passed by pointer, so the arrays are pinned for the
	// duration of the call (runtime.Pinner, Go >= 1.21; go.mod says 1.23).
	t := C.frc_tree_t{n_nodes: C.int32_t(len(parent)), parent: &parent[0], length: &length[0]}
	a := C.frc_csr_t{n_samples: C.int64_t(len(abnd)), row_ptr: &rowPtr[0]}
	if len(col) > 0 {
		a.col, a.val = &col[0], &val[0]
	}
	pinInputs := func(p *runtime.Pinner) {
		p.Pin(&parent[0])
		p.Pin(&length[0])
		p.Pin(&rowPtr[0])
		if len(col) > 0 {
			p.Pin(&col[0])
			p.Pin(&val[0])
		}
	}
	// n_devices = -1: every visible B200 works on the job; the engine validates and stages the inputs once and
	// hands the bands of all devices out as ONE ordered stream, so this function stays the single iter.Seq the
	// caller ranges over (frcfrc.go:58-62).
	o := C.frc_opts_t{mode: C.FRC_UNWEIGHTED, normalize: 1, path: C.FRC_PATH_AUTO, device: -1, world: 1, n_devices: -1}
	if weighted {
		o.mode = C.FRC_WEIGHTED
	}
	if *nnorm { // flag -l (frcfrc.go:25): 0 = what unifrac.go:108-110 does (normalizeFlatNodes skipped, and with it
		o.normalize = 0 // the id-sort of :57); 2 would be the documented behaviour (sorted lists, raw values)
	}
	// *nt (flag -p) has no meaning on the device; it keeps sizing the parsers' goroutine pools.

	return func(yield func(float64) bool) {
		// frc_last_error(NULL) reads a thread-local of the C library: the create call and the read of its
		// message must run on the same OS thread
		runtime.LockOSThread()
		var job *C.frc_job_t
		var pin runtime.Pinner
		pinInputs(&pin)
		rc := C.frc_create(nil, &t, &a, &o, &job)
		pin.Unpin() // everything was copied
		if rc != C.FRC_OK {
			msg := C.GoString(C.frc_last_error(nil))
			runtime.UnlockOSThread()
			fmt.Fprintln(os.Stderr, "ERROR:", msg) // common.ExitIfError
			os.Exit(2)
		}
		runtime.UnlockOSThread()
		defer C.frc_destroy(job) // legal mid-stream: the consumer's `break` (frcfrc.go:59-61)
		fmt.Fprintln(os.Stderr, "Calculating distances")
		var info C.frc_info_t
		C.frc_job_info(job, &info)
		fail := func() {
			fmt.Fprintln(os.Stderr, "ERROR:", C.GoString(C.frc_last_error(job)))
			os.Exit(2)
		}
		if info.value_bytes == 4 {
			// fast paths: the kernels produce fp32 ratios; take the pinned fp32 bands as they are and widen
			// per value here (float64(f) is exact), instead of a widening pass over memory inside the library
			for {
				var data *C.float
				var first, n C.int64_t
				if rc := C.frc_next_f32(job, &data, &first, &n); rc != C.FRC_OK {
					fail()
				}
				if n == 0 {
					return
				}
				band := unsafe.Slice((*float32)(unsafe.Pointer(data)), int(n))
				// distances below fp32's range (pathological trees only) travel beside the band as float64
				var xi *C.int64_t
				var xv *C.double
				var nx C.int64_t
				C.frc_chunk_exceptions(job, &xi, &xv, &nx)
				patch := map[int]float64{}
				if nx > 0 {
					idx := unsafe.Slice((*int64)(unsafe.Pointer(xi)), int(nx))
					val := unsafe.Slice((*float64)(unsafe.Pointer(xv)), int(nx))
					for k := range idx {
						patch[int(idx[k]-int64(first))] = val[k]
					}
				}
				for k, f := range band {
					v := float64(f)
					if nx > 0 {
						if p, ok := patch[k]; ok {
							v = p
						}
					}
					if !yield(v) { // pinned memory stays valid until the next frc_next_f32
						return
					}
				}
			}
		}
		for { // exact path: bit-exact float64
			var data *C.double
			var first, n C.int64_t
			if rc := C.frc_next(job, &data, &first, &n); rc != C.FRC_OK {
				fail()
			}
			if n == 0 {
				return
			}
			for _, f := range unsafe.Slice((*float64)(unsafe.Pointer(data)), int(n)) {
				if !yield(f) { // pinned memory stays valid until the next frc_next
					return
				}
			}
		}
	}
}
