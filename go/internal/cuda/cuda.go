// Package cuda is the cgo binding that replaces internal/ocl's OpenCL driver in
// eriklupander/pathtracer-ocl with libptcuda (hand-written sm_100a kernels behind a C ABI).
//
// It is a drop-in for the two things the rest of the Go program uses from the OpenCL side:
//
//	ocl.Trace(objects, triangles, groups, deviceIndex, samples, camera, textures, sphereTextures, cubeTextures) []float64
//	    (internal/ocl/ocltracer.go:100)  ->  cuda.Trace(... same arguments ...)
//	listDevices()  (cmd/pt/main.go:98-112)  ->  cuda.ListDevices()
//
// The scene records (CLObject / CLTriangle / CLGroup / CLCamera, ocltracer.go:25-96) are passed through
// untouched: libptcuda consumes the same packed bytes the OpenCL kernel did (include/ptwire.h).  This package
// does NOT import internal/ocl -- that package's ocltracer.go pulls in github.com/jgillich/go-opencl/cl (cgo,
// OpenCL headers).  Trace is generic over the record types instead and checks their sizes (1024 / 512 / 256 /
// 256 bytes) at the call, so the call site in renderer.go compiles unchanged with the structs wherever they
// live, and internal/cuda itself builds with no OpenCL header, ICD or runtime present.
//
// NOTE: no Go toolchain exists in the image this was written in.  The C side of every call below is exercised,
// with exactly these argument shapes, by tests/c_abi_smoke.c; the Go side has been checked by eye only.  The
// generic functions contain no cgo calls (they only check sizes and take addresses; traceRaw does the rest).
package cuda

/*
#cgo CFLAGS: -I${SRCDIR}/../../../include
#cgo LDFLAGS: -L${SRCDIR}/../../../pathtracer_ocl_b200 -lptcuda -Wl,-rpath,${SRCDIR}/../../../pathtracer_ocl_b200
#include <stdlib.h>
#include <string.h>
#include "ptcuda.h"
*/
import "C"

import (
	"fmt"
	"image"
	"math/rand"
	"unsafe"

	"github.com/sirupsen/logrus"
)

// Wire record sizes (include/ptwire.h; ocltracer.go:25-96).
const (
	objectBytes   = 1024
	triangleBytes = 512
	groupBytes    = 256
	cameraBytes   = 256
)

// Options selects what the OpenCL path could not: arithmetic precision, RNG evaluation, several GPUs,
// and the two code paths upstream ships disabled.
type Options struct {
	FP64         bool    // false: fp32 mode (default), true: fp64, the precision of tracer.cl
	FastRNG      bool    // cheaper evaluation of the same noise3D hash (still reproducible)
	Devices      []int32 // GPUs driven by this process; nil = {deviceIndex}
	Seeds        []float64
	NEE          bool // next-event estimation (tracer.cl:786-825, call commented out at :1168)
	CylinderCaps bool // cylinder end caps (tracer.cl:282-310, disabled at :437-444)
}

// ListDevices prints what the reference's --list-devices prints (cmd/pt/main.go:98-112).
func ListDevices() {
	n := int(C.ptc_device_count())
	for i := 0; i < n; i++ {
		buf := make([]byte, 256)
		if C.ptc_device_name(C.int(i), (*C.char)(unsafe.Pointer(&buf[0])), C.int(len(buf))) == 0 {
			fmt.Printf("Index: %d Type: GPU Name: %s\n", i, C.GoString((*C.char)(unsafe.Pointer(&buf[0]))))
		}
	}
}

// Trace has the shape of ocl.Trace (internal/ocl/ocltracer.go:100): called with []CLObject, []CLTriangle,
// []CLGroup and a CLCamera it needs no type arguments at the call site.
func Trace[O, T, G, Cam any](objects []O, triangles []T, groups []G, deviceIndex, samples int,
	camera Cam, textures []image.Image, sphereTextures []image.Image, cubeTextures []image.Image) []float64 {
	return TraceWithOptions(objects, triangles, groups, deviceIndex, samples, camera, textures, sphereTextures, cubeTextures, Options{})
}

// packTextures mirrors prepareTextures (ocltracer.go:228-254): same-sized *image.NRGBA packed back to back.
func packTextures(imgs []image.Image) ([]byte, int, int, int) {
	if len(imgs) == 0 {
		return nil, 0, 0, 0
	}
	first := imgs[0].(*image.NRGBA)
	w, h := first.Bounds().Dx(), first.Bounds().Dy()
	all := make([]byte, 0, len(imgs)*w*h*4)
	for _, im := range imgs {
		n := im.(*image.NRGBA)
		if n.Bounds().Dx() != w || n.Bounds().Dy() != h {
			logrus.Fatalf("textures of one class must share a size")
		}
		if n.Stride == w*4 {
			all = append(all, n.Pix[:w*h*4]...)
		} else {
			for y := 0; y < h; y++ {
				all = append(all, n.Pix[y*n.Stride:y*n.Stride+w*4]...)
			}
		}
	}
	return all, w, h, len(imgs)
}

func checkSize[R any](what string, want uintptr) {
	var zero R
	if unsafe.Sizeof(zero) != want {
		logrus.Fatalf("cuda.Trace: %s records are %d bytes, the wire format has %d", what, unsafe.Sizeof(zero), want)
	}
}

// TraceWithOptions is Trace plus what the OpenCL path could not select.  The generic part only checks the record sizes
// and takes addresses; every cgo call lives in the non-generic traceRaw below.
func TraceWithOptions[O, T, G, Cam any](objects []O, triangles []T, groups []G, deviceIndex, samples int,
	camera Cam, textures []image.Image, sphereTextures []image.Image, cubeTextures []image.Image, opt Options) []float64 {

	checkSize[O]("object", objectBytes)
	checkSize[T]("triangle", triangleBytes)
	checkSize[G]("group", groupBytes)
	checkSize[Cam]("camera", cameraBytes)
	if len(objects) == 0 {
		logrus.Fatalf("cuda.Trace: scene has no objects")
	}
	var triPtr, grpPtr unsafe.Pointer
	if len(triangles) > 0 {
		triPtr = unsafe.Pointer(&triangles[0])
	}
	if len(groups) > 0 {
		grpPtr = unsafe.Pointer(&groups[0])
	}
	return traceRaw(unsafe.Pointer(&objects[0]), len(objects), triPtr, len(triangles), grpPtr, len(groups),
		unsafe.Pointer(&camera), deviceIndex, samples, textures, sphereTextures, cubeTextures, opt)
}

// traceRaw passes the packed records (include/ptwire.h) to ptc_render_flat2.  camera points at a 256-byte CLCamera.
func traceRaw(objects unsafe.Pointer, nObjects int, triangles unsafe.Pointer, nTriangles int, groups unsafe.Pointer, nGroups int,
	camera unsafe.Pointer, deviceIndex, samples int, textures, sphereTextures, cubeTextures []image.Image, opt Options) []float64 {

	// CLCamera starts with Width, Height int32 (ocltracer.go:86-87)
	dims := (*[2]int32)(camera)
	width, height := int(dims[0]), int(dims[1])
	numPixels := width * height
	logrus.Infof("trace with %d objects %dx%d", nObjects, width, height)

	// one random double per pixel, as computeBatch does per batch (ocltracer.go:260-263)
	seeds := opt.Seeds
	if seeds == nil {
		seeds = make([]float64, numPixels)
		for i := range seeds {
			seeds[i] = rand.Float64()
		}
	}

	// Every buffer goes to C as a direct call argument (ptc_render_flat2): cgo pins Go memory passed
	// that way for the duration of the call, whereas storing Go pointers inside a C struct is not allowed.
	var tex [3]*C.uint8_t
	var texDims [9]C.int32_t
	var keep [3][]byte
	for cls, imgs := range [][]image.Image{textures, sphereTextures, cubeTextures} {
		pix, w, h, layers := packTextures(imgs)
		if layers == 0 {
			continue
		}
		keep[cls] = pix
		tex[cls] = (*C.uint8_t)(unsafe.Pointer(&pix[0]))
		texDims[3*cls], texDims[3*cls+1], texDims[3*cls+2] = C.int32_t(w), C.int32_t(h), C.int32_t(layers)
	}
	precision, rngMode := C.int32_t(C.PTC_FP32), C.int32_t(C.PTC_RNG_PARITY)
	if opt.FP64 {
		precision = C.PTC_FP64
	}
	if opt.FastRNG {
		rngMode = C.PTC_RNG_FAST
	}
	features := C.int32_t(0)
	if opt.NEE {
		features |= C.PTC_FEATURE_NEE
	}
	if opt.CylinderCaps {
		features |= C.PTC_FEATURE_CYLINDER_CAPS
	}
	devices := opt.Devices
	if devices == nil {
		devices = []int32{int32(deviceIndex)}
	}

	results := make([]float64, numPixels*4)
	errbuf := make([]byte, 512)
	rc := C.ptc_render_flat2(objects, C.int32_t(nObjects), triangles, C.int32_t(nTriangles),
		groups, C.int32_t(nGroups), camera, tex[0], tex[1], tex[2], &texDims[0],
		(*C.double)(unsafe.Pointer(&seeds[0])), C.int32_t(samples), precision, rngMode, features,
		(*C.int32_t)(unsafe.Pointer(&devices[0])), C.int32_t(len(devices)),
		(*C.double)(unsafe.Pointer(&results[0])), (*C.char)(unsafe.Pointer(&errbuf[0])), C.int(len(errbuf)))
	if rc != 0 {
		logrus.Fatalf("ptc_render failed: %s", C.GoString((*C.char)(unsafe.Pointer(&errbuf[0]))))
	}
	_ = keep
	return results
}
