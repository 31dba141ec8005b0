// Package coregex — cgo binding that routes the bulk-scan methods of coregx/coregex to the B200
// engine (libcoregex_b200.so).  Source only in this repository: the build image has no Go
// toolchain, so this file is compiled where Go >= 1.22 and the shared library are available:
//
//	CGO_CFLAGS="-I${REPO}/include" CGO_LDFLAGS="-L${REPO}/coregex_b200/lib -lcoregex_b200" go build ./go/coregex
//
// The method set and semantics are those of reference regex.go: Compile :110, MustCompile :129,
// Match :282, FindAllIndex :695, Count :1349, FindAllSubmatchIndex :1423, NumSubexp :552,
// String :444.  cgo rules honoured: the haystack slice is pinned for the duration of the call
// (runtime.Pinner) and never retained by C; results are written into Go-owned memory.
package coregex

/*
#cgo LDFLAGS: -lcoregex_b200
#include <stdlib.h>
#include "coregex_b200.h"
*/
import "C"

import (
	"errors"
	"runtime"
	"unsafe"
)

// Regex mirrors reference regex.go `type Regex` for the bulk-scan path.
type Regex struct {
	h       *C.cgx_regex
	pattern string
}

// Regexp is the stdlib-compatible alias (reference regex.go:97).
type Regexp = Regex

// Compile mirrors reference regex.go:110. Syntax errors carry the stdlib text
// ("error parsing regexp: ...", reference meta/compile.go:775-784).
func Compile(pattern string) (*Regex, error) {
	cp := C.CString(pattern)
	defer C.free(unsafe.Pointer(cp))
	var h *C.cgx_regex
	errbuf := make([]byte, 1024)
	rc := C.cgx_compile(cp, C.size_t(len(pattern)), &h, (*C.char)(unsafe.Pointer(&errbuf[0])), C.size_t(len(errbuf)))
	if rc != C.CGX_OK {
		n := 0
		for n < len(errbuf) && errbuf[n] != 0 {
			n++
		}
		return nil, errors.New(string(errbuf[:n]))
	}
	re := &Regex{h: h, pattern: pattern}
	runtime.SetFinalizer(re, func(r *Regex) { C.cgx_free(r.h) })
	return re, nil
}

// MustCompile mirrors reference regex.go:129 (panic text built at :132).
func MustCompile(pattern string) *Regex {
	re, err := Compile(pattern)
	if err != nil {
		panic("regexp: Compile(`" + pattern + "`): " + err.Error())
	}
	return re
}

func (r *Regex) String() string { return r.pattern }                      // regex.go:444
func (r *Regex) NumSubexp() int { return int(C.cgx_num_captures(r.h)) - 1 } // regex.go:552
func (r *Regex) Strategy() string { return C.GoString(C.cgx_strategy(r.h)) }

func ptr(b []byte) *C.uint8_t {
	if len(b) == 0 {
		return nil
	}
	return (*C.uint8_t)(unsafe.Pointer(&b[0]))
}

// Match mirrors reference regex.go:282 (meta.Engine.IsMatch, meta/ismatch.go:27).
func (r *Regex) Match(b []byte) bool {
	var p runtime.Pinner
	if len(b) > 0 {
		p.Pin(&b[0])
		defer p.Unpin()
	}
	var m C.int
	if C.cgx_is_match(r.h, ptr(b), C.size_t(len(b)), &m) != C.CGX_OK {
		panic("coregex_b200: " + C.GoString(C.cgx_last_error()))
	}
	return m != 0
}

// Count mirrors reference regex.go:1349 (meta.Engine.Count, meta/findall.go:297).
func (r *Regex) Count(b []byte, n int) int {
	var p runtime.Pinner
	if len(b) > 0 {
		p.Pin(&b[0])
		defer p.Unpin()
	}
	var c C.size_t
	if C.cgx_count(r.h, ptr(b), C.size_t(len(b)), C.int64_t(n), &c) != C.CGX_OK {
		panic("coregex_b200: " + C.GoString(C.cgx_last_error()))
	}
	return int(c)
}

// AppendAllIndex mirrors reference regex.go:748: reuses dst[:0], grows on demand (two-call sizing).
func (r *Regex) AppendAllIndex(dst [][2]int, b []byte, n int) [][2]int {
	dst = dst[:0]
	if n == 0 {
		return dst
	}
	var p runtime.Pinner
	if len(b) > 0 {
		p.Pin(&b[0])
		defer p.Unpin()
	}
	capPairs := cap(dst)
	if capPairs < 256 {
		capPairs = len(b)/100 + 256
	}
	for {
		buf := make([][2]int, capPairs) // Go int is 64-bit: same layout as int64 pairs
		var c C.size_t
		rc := C.cgx_find_all_index(r.h, ptr(b), C.size_t(len(b)), C.int64_t(n),
			(*C.int64_t)(unsafe.Pointer(&buf[0])), C.size_t(capPairs), &c)
		if rc != C.CGX_OK {
			panic("coregex_b200: " + C.GoString(C.cgx_last_error()))
		}
		if int(c) <= capPairs {
			return buf[:int(c)]
		}
		capPairs = int(c)
	}
}

// FindAllIndex mirrors reference regex.go:695-723: nil when n == 0 or nothing matches; each inner
// slice has len 2 / cap 2 over one flat backing array.
func (r *Regex) FindAllIndex(b []byte, n int) [][]int {
	pairs := r.AppendAllIndex(nil, b, n)
	if len(pairs) == 0 {
		return nil
	}
	flat := make([]int, 2*len(pairs))
	out := make([][]int, len(pairs))
	for i, m := range pairs {
		flat[2*i], flat[2*i+1] = m[0], m[1]
		out[i] = flat[2*i : 2*i+2 : 2*i+2]
	}
	return out
}

// FindAllStringIndex mirrors reference regex.go:777 (zero-copy view of the string bytes).
func (r *Regex) FindAllStringIndex(s string, n int) [][]int {
	return r.FindAllIndex(unsafe.Slice(unsafe.StringData(s), len(s)), n)
}

// FindAllSubmatchIndex mirrors reference regex.go:1423: stride 2*(NumSubexp()+1), -1 for unmatched groups.
func (r *Regex) FindAllSubmatchIndex(b []byte, n int) [][]int {
	if n == 0 {
		return nil
	}
	var p runtime.Pinner
	if len(b) > 0 {
		p.Pin(&b[0])
		defer p.Unpin()
	}
	stride := 2 * (r.NumSubexp() + 1)
	capM := len(b)/64 + 256
	for {
		flat := make([]int, capM*stride)
		var c C.size_t
		rc := C.cgx_find_all_submatch_index(r.h, ptr(b), C.size_t(len(b)), C.int64_t(n),
			(*C.int64_t)(unsafe.Pointer(&flat[0])), C.size_t(capM), &c)
		if rc != C.CGX_OK {
			panic("coregex_b200: " + C.GoString(C.cgx_last_error()))
		}
		if int(c) <= capM {
			if c == 0 {
				return nil
			}
			out := make([][]int, int(c))
			for i := range out {
				out[i] = flat[i*stride : (i+1)*stride : (i+1)*stride]
			}
			return out
		}
		capM = int(c)
	}
}

// SubexpNames mirrors reference regex.go:575 (names[0] == "").
func (r *Regex) SubexpNames() []string {
	names := make([]string, r.NumSubexp()+1)
	for i := range names {
		names[i] = C.GoString(C.cgx_subexp_name(r.h, C.int(i)))
	}
	return names
}

// ReplaceAllLiteral mirrors reference regex.go:790.  The reference walks the haystack match by match
// (FindIndicesAt from the previous end, :797-840); that walk IS the FindAllIndex loop, so here it is
// one batch call and a stitch.
func (r *Regex) ReplaceAllLiteral(src, repl []byte) []byte {
	pairs := r.AppendAllIndex(nil, src, -1)
	out := make([]byte, 0, len(src))
	last := 0
	for _, m := range pairs {
		out = append(out, src[last:m[0]]...)
		out = append(out, repl...)
		last = m[1]
	}
	return append(out, src[last:]...)
}

// Split mirrors reference regex.go:1288 (same skips of an empty match at 0 and at len(s)).
func (r *Regex) Split(s string, n int) []string {
	if n == 0 {
		return nil
	}
	idx := r.FindAllStringIndex(s, -1)
	if len(idx) == 0 {
		return []string{s}
	}
	var out []string
	last := 0
	for _, m := range idx {
		if last == 0 && m[0] == 0 && m[1] == 0 {
			continue
		}
		if m[0] == len(s) && m[1] == len(s) {
			break
		}
		out = append(out, s[last:m[0]])
		last = m[1]
		if n > 0 && len(out) >= n-1 {
			return append(out, s[last:])
		}
	}
	return append(out, s[last:])
}
