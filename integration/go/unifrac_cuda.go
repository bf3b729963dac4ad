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
	o := C.frc_opts_t{mode: C.FRC_UNWEIGHTED, normalize: 1, path: C.FRC_PATH_AUTO, device: -1, world: 1}
	if weighted {
		o.mode = C.FRC_WEIGHTED
	}
	if *nnorm { // flag -l (frcfrc.go:25)
		o.normalize = 0
	}
	// *nt (flag -p) has no meaning on the device; it keeps sizing the parsers' goroutine pools.

	return func(yield func(float64) bool) {
		var job *C.frc_job_t
		var pin runtime.Pinner
		pinInputs(&pin)
		rc := C.frc_create(nil, &t, &a, &o, &job)
		pin.Unpin() // everything was copied
		if rc != C.FRC_OK {
			fmt.Fprintln(os.Stderr, "ERROR:", C.GoString(C.frc_last_error(nil))) // common.ExitIfError
			os.Exit(2)
		}
		defer C.frc_destroy(job) // legal mid-stream: the consumer's `break` (frcfrc.go:59-61)
		fmt.Fprintln(os.Stderr, "Calculating distances")
		for {
			var data *C.double
			var first, n C.int64_t
			if rc := C.frc_next(job, &data, &first, &n); rc != C.FRC_OK {
				fmt.Fprintln(os.Stderr, "ERROR:", C.GoString(C.frc_last_error(job)))
				os.Exit(2)
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
