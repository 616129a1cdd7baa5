// Package inference is the cgo shim between ACP's llmclient and libacp_infer.so
// (include/acp_infer.h).  It would live at acp/internal/inference in the reference tree.
//
// NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no Go toolchain.  Every call below
// has a one-to-one C++ twin that IS compiled and tested here
// (agentcontrolplane_b200/csrc/host/llmclient.cc, LocalClient::SendRequest).
//
// Threading: a blocking cgo call pins an OS thread, so 1000 concurrent SendRequest goroutines
// must not each sit in acp_infer_wait.  Submit is non-blocking; ONE poller goroutine sits in
// acp_infer_poll and fans completions out to per-ticket channels.
package inference

/*
#cgo CFLAGS: -I${SRCDIR}/../../../include
#cgo LDFLAGS: -L${SRCDIR}/../../../agentcontrolplane_b200/lib -lacp_infer
#include <stdlib.h>
#include "acp_infer.h"
*/
import "C"

import (
	"context"
	"fmt"
	"sync"
	"time"
	"unsafe"
)

// Engine is the process-wide singleton created once in cmd/main.go right after the
// TaskReconciler is wired (acp/cmd/main.go:253-261).
type Engine struct {
	h       *C.acp_engine
	mu      sync.Mutex
	waiters map[uint64]chan struct{}
	done    chan struct{}
	wg      sync.WaitGroup // the poller; Shutdown waits for it before the handle is freed
	closed  bool
}

var (
	global     *Engine
	globalOnce sync.Once
	globalErr  error
)

// Init starts the engine (weights -> HBM, KV pool, scheduler thread).  configJSON is the JSON
// form of LLMSpec.Local (see INTEGRATION.md).
func Init(configJSON string) (*Engine, error) {
	globalOnce.Do(func() {
		cs := C.CString(configJSON)
		defer C.free(unsafe.Pointer(cs))
		var h *C.acp_engine
		if rc := C.acp_infer_init(cs, &h); rc != C.ACP_OK {
			globalErr = fmt.Errorf("acp_infer_init failed: %d", int(rc))
			return
		}
		global = &Engine{h: h, waiters: map[uint64]chan struct{}{}, done: make(chan struct{})}
		global.wg.Add(1)
		go global.poller()
	})
	return global, globalErr
}

// Get returns the singleton or nil when Init has not run.
func Get() *Engine { return global }

func (e *Engine) poller() {
	defer e.wg.Done()
	var tickets [256]C.uint64_t
	for {
		select {
		case <-e.done:
			return
		default:
		}
		n := int(C.acp_infer_poll(e.h, &tickets[0], 256, 100))
		if n <= 0 {
			continue
		}
		e.mu.Lock()
		for i := 0; i < n; i++ {
			if ch, ok := e.waiters[uint64(tickets[i])]; ok {
				close(ch)
				delete(e.waiters, uint64(tickets[i]))
			}
		}
		e.mu.Unlock()
	}
}

// Complete submits one OpenAI chat-completions body and blocks the calling goroutine (not an OS
// thread) until the engine finishes it or ctx is cancelled.
func (e *Engine) Complete(ctx context.Context, body []byte) (status int, resp []byte, err error) {
	var ticket C.uint64_t
	ch := make(chan struct{})
	e.mu.Lock() // register before submit so the poller cannot miss the completion
	if e.closed {
		e.mu.Unlock()
		return 0, nil, fmt.Errorf("engine is shut down")
	}
	rc := C.acp_infer_submit(e.h, (*C.char)(unsafe.Pointer(&body[0])), C.size_t(len(body)), &ticket)
	if rc != C.ACP_OK {
		e.mu.Unlock()
		return 0, nil, fmt.Errorf("acp_infer_submit failed: %d", int(rc))
	}
	e.waiters[uint64(ticket)] = ch
	e.mu.Unlock()

	select {
	case <-ch:
	case <-ctx.Done(): // manager shutdown: drop the sequence, then collect its 499
		C.acp_infer_cancel(e.h, ticket)
		<-ch
	}
	var out *C.char
	var n C.size_t
	var st C.int
	if rc := C.acp_infer_result(e.h, ticket, &out, &n, &st); rc != C.ACP_OK {
		return 0, nil, fmt.Errorf("acp_infer_result failed: %d", int(rc))
	}
	defer C.acp_infer_free(unsafe.Pointer(out))
	return int(st), C.GoBytes(unsafe.Pointer(out), C.int(n)), nil
}

// Shutdown stops the scheduler and frees device memory.  acp_infer_shutdown deletes the handle, so
// nothing may be inside acp_infer_poll / acp_infer_result when it runs (include/acp_infer.h:
// "shutdown must not overlap other calls"):
//  1. refuse new submissions, cancel what is in flight (each ends with status 499 and is reported
//     to the poller like any completion, so parked Complete() goroutines wake up);
//  2. stop the poller and WAIT for it to leave acp_infer_poll;
//  3. wake any goroutine still parked (its acp_infer_result then reports the engine's 503 / 499);
//  4. free the handle.
func (e *Engine) Shutdown() {
	e.mu.Lock()
	if e.closed {
		e.mu.Unlock()
		return
	}
	e.closed = true
	for t := range e.waiters {
		C.acp_infer_cancel(e.h, C.uint64_t(t))
	}
	e.mu.Unlock()
	// give the scheduler one poll interval to report the cancellations through the poller
	drained := make(chan struct{})
	go func() {
		for {
			e.mu.Lock()
			n := len(e.waiters)
			e.mu.Unlock()
			if n == 0 {
				close(drained)
				return
			}
			select {
			case <-e.done:
				return
			case <-time.After(10 * time.Millisecond):
			}
		}
	}()
	select {
	case <-drained:
	case <-time.After(2 * time.Second):
	}
	close(e.done)
	e.wg.Wait()
	e.mu.Lock()
	for t, ch := range e.waiters {
		close(ch)
		delete(e.waiters, t)
	}
	e.mu.Unlock()
	C.acp_infer_shutdown(e.h)
}
